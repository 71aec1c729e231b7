"""Print the gather modes (dim, stride in A, B, C) of the heavy steps of a bench workload's
plan -- runs on the CPU (planning needs no GPU).  python tools/dump_steps.py cfg4 [min_macs]"""
import os
import pickle
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from optyx_b200 import executor  # noqa: E402
from optyx_b200.core.backends import B200Backend  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
w = bench.workload(name)
kw = dict(target_size=w["target"], trials=w["trials"], min_slices=w.get("min_slices", 1))
net = B200Backend(**kw)._get_network(w["factory"]())
plan = executor.get_plan(net, **kw)
rows = []
for k, st in enumerate(plan.steps):
    if st.grouped or st.macs < float(sys.argv[2] if len(sys.argv) > 2 else 1e6):
        continue
    d = st.desc
    if st.kernel == "gemm":
        n = d.n_batch + d.n_m + d.n_n + d.n_k
        groups = ["B"] * d.n_batch + ["M"] * d.n_m + ["N"] * d.n_n + ["K"] * d.n_k
        modes = [(groups[i], d.dim[i], d.sa[i], d.sb[i], d.sc[i]) for i in range(n)]
    else:
        modes = "direct"
    rows.append((k, st.kernel, st.mnk, st.per_slice, modes))
    print(k, st.kernel, st.mnk, "S" if st.per_slice else "-", modes, flush=True)
with open("/tmp/steps_%s.pkl" % name, "wb") as fh:
    pickle.dump(rows, fh)
