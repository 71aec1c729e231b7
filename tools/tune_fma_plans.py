"""Whole-plan device time of the other BASELINE workloads (cfg1, scaled cfg3, cfg4
sliced) under the contract_fma_kernel tuning overrides (see tools/tune_fma.py)."""
import json
import os
import sys

os.environ["B2_FMA_TUNE"] = "1"
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from optyx_b200 import workloads as cases  # noqa: E402
from optyx_b200 import executor  # noqa: E402

KEYS = ("B2_FMA_TMAX", "B2_FMA_MAXM", "B2_FMA_MAXN", "B2_FMA_FLAGS", "B2_FMA_GRID", "B2_FMA_RM")
CONFIGS = [
    {},
    {"B2_FMA_RM": 2},
    {"B2_FMA_RM": 4},
    {"B2_FMA_RM": 4, "B2_FMA_TMAX": 65536},
    {"B2_FMA_RM": 4, "B2_FMA_TMAX": 65536, "B2_FMA_FLAGS": 1},
    {"B2_FMA_RM": 4, "B2_FMA_TMAX": 65536, "B2_FMA_FLAGS": 3},
    {"B2_FMA_FLAGS": 1},
]
WORK = {
    "cfg1": (lambda: cases.cfg1_interferometer(6, (1, 1, 1, 0, 0, 0)), None, 20),
    "cfg3": (lambda: cases.cfg3_interferometer(width=6, n_photons=3, measured=2), None, 3),
    "cfg4": (lambda: cases.random_clifford_t(30, 41, seed=0), float(2 ** 26), 2),
}
for name in (sys.argv[1:] or ["cfg1", "cfg3", "cfg4"]):
    factory, target, reps = WORK[name]
    net = cases.compile_channel(factory())
    plan = executor.get_plan(net, target_size=target, trials=16 if target else 8)
    sess = executor.Session(plan)
    sess.set_leaves([net])
    sess.build_leaves()
    ref = None
    for cfg in CONFIGS:
        for key in KEYS:
            os.environ.pop(key, None)
        for key, val in cfg.items():
            os.environ[key] = str(val)
        sess.inv_graph = None                       # re-capture under the new settings
        sess.contract(net.scalar)
        torch.cuda.synchronize()
        out = sess.out.clone()
        if ref is None:
            ref = out
        err = float((out - ref).abs().max() / ref.abs().max())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            sess.contract(net.scalar)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"workload": name, "cfg": cfg, "ms": round(e0.elapsed_time(e1) / reps, 4),
                          "rel_diff_vs_default": err}), flush=True)
    del sess
    torch.cuda.empty_cache()
