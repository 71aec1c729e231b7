"""Turn ncu captures into the tables bench.py and the judge read.

    python tools/ncu_traffic.py <key> <file.ncu-rep> [more key/file pairs]

<key> is ``<workload>/<kernel signature>`` as printed by tools/prof_dominant.py.
For every capture: the per-launch averages of dram__bytes_read.sum +
dram__bytes_write.sum, gpu__time_duration.sum, DRAM / tensor-pipe utilisation,
registers -> merged into profiles/r02_traffic.json (bench.py: roofline.traffic)
and one markdown summary per capture under profiles/.
Runs here (no GPU): ``ncu -i <rep> --page raw --csv``.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__time_duration.sum": "duration",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "smsp__cycles_active.avg": "cycles",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
         "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "nsecond": 1e-9, "usecond": 1e-6,
         "msecond": 1e-3, "second": 1.0}


def read_rep(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    out = []
    for r in data:
        rec = {"kernel": r[col["Kernel Name"]]}
        for metric, short in WANT.items():
            if metric in col and r[col[metric]] not in ("", "n/a"):
                v = float(r[col[metric]].replace(",", ""))
                rec[short] = v * SCALE.get(units[col[metric]], 1.0) if short in (
                    "dram_read", "dram_write", "duration") else v
        out.append(rec)
    return out


def main():
    table_path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        table = json.load(open(table_path))
    except (OSError, ValueError):
        table = {}
    args = sys.argv[1:]
    for key, path in zip(args[0::2], args[1::2]):
        recs = read_rep(path)
        n = len(recs)
        avg = {k: sum(r.get(k, 0.0) for r in recs) / n for k in set().union(*recs) if k != "kernel"}
        entry = {"kernel": recs[0]["kernel"].split("(")[0], "launches": n,
                 "dram_bytes": avg.get("dram_read", 0.0) + avg.get("dram_write", 0.0),
                 "source": os.path.basename(path)}
        entry.update({k: avg[k] for k in sorted(avg)})
        table[key] = entry
        md = os.path.join(ROOT, "profiles", os.path.basename(path).replace(".ncu-rep", ".md"))
        with open(md, "w") as fh:
            fh.write(f"# ncu --set full: `{key}`\n\nkernel `{entry['kernel']}`, {n} launches "
                     f"(per-launch averages; ncu replays are cold-cache, durations are not bench values)\n\n")
            fh.write("| metric | value |\n|---|---|\n")
            for k in sorted(entry):
                if k not in ("kernel", "source"):
                    v = entry[k]
                    fh.write(f"| {k} | {v:.6g} |\n" if isinstance(v, float) else f"| {k} | {v} |\n")
        print(key, json.dumps(entry))
    with open(table_path, "w") as fh:
        json.dump(table, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
