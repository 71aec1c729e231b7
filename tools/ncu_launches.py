"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel:
launch count, summed device time, share.  python tools/ncu_launches.py in.csv out.md "title" """
import csv
import sys
from collections import OrderedDict

src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = []
with open(src, newline="") as fh:
    lines = [l for l in fh if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        rows.append((r["Kernel Name"].split("(")[0], v))
agg = OrderedDict()
for k, v in rows:
    a = agg.setdefault(k, [0, 0.0, 0.0])
    a[0] += 1; a[1] += v; a[2] = max(a[2], v)
tot = sum(v for _, v in rows)
with open(dst, "w") as fh:
    fh.write(f"# {title}\n\n{len(rows)} launches, {tot / 1e3:.3f} ms summed device time "
             "(ncu serialises and runs every launch cold-cache: compare SHARES, not absolutes)\n\n")
    fh.write("| kernel | launches | summed us | share | longest us |\n|---|---|---|---|---|\n")
    for k, (n, t, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f} % | {mx:.1f} |\n")
print(open(dst).read())
