"""Sweep the tile limits / flags of contract_fma_kernel on the dominant (final) step of
the cfg2 bench workload.  B2_FMA_TUNE=1 makes the library re-read its B2_FMA_*
environment overrides on every launch, so one process covers the whole sweep.
usage: python tools/tune_fma.py [k_open]   (prints one JSON line per configuration)"""
import json
import os
import sys

os.environ["B2_FMA_TUNE"] = "1"
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from optyx_b200 import workloads as cases  # noqa: E402
from optyx_b200 import executor, runtime  # noqa: E402

KEYS = ("B2_FMA_TMAX", "B2_FMA_MAXM", "B2_FMA_MAXN", "B2_FMA_FLAGS", "B2_FMA_GRID", "B2_FMA_RM")
CONFIGS = [
    {},
    {"B2_FMA_RM": 4, "B2_FMA_TMAX": 65536, "B2_FMA_FLAGS": 1},
    {"B2_FMA_RM": 4, "B2_FMA_TMAX": 65536, "B2_FMA_FLAGS": 3},
    {"B2_FMA_RM": 2, "B2_FMA_TMAX": 65536, "B2_FMA_FLAGS": 1},
    {"B2_FMA_RM": 2, "B2_FMA_TMAX": 65536, "B2_FMA_FLAGS": 3},
    {"B2_FMA_RM": 4, "B2_FMA_TMAX": 32768, "B2_FMA_MAXM": 512, "B2_FMA_FLAGS": 1},
    {"B2_FMA_RM": 4, "B2_FMA_TMAX": 32768, "B2_FMA_MAXM": 512, "B2_FMA_FLAGS": 3},
    {"B2_FMA_RM": 2, "B2_FMA_TMAX": 32768, "B2_FMA_MAXM": 512, "B2_FMA_FLAGS": 3},
    {"B2_FMA_RM": 4, "B2_FMA_TMAX": 32768, "B2_FMA_MAXN": 32, "B2_FMA_FLAGS": 3},
    {"B2_FMA_RM": 4, "B2_FMA_TMAX": 65536, "B2_FMA_FLAGS": 3, "B2_FMA_GRID": 1},
]

k = int(sys.argv[1]) if len(sys.argv) > 1 else 14
net = cases.compile_channel(cases.ghz(20, open_qubits=k, noise=0.95))
plan = executor.get_plan(net)
sess = executor.Session(plan)
sess.run([net])
torch.cuda.synchronize()
ref = sess.out.clone()
dom = max(plan.steps, key=lambda s: s.bytes)
stream = torch.cuda.current_stream().cuda_stream
lib = runtime.load()
reps = 10
for cfg in CONFIGS:
    for key in KEYS:
        os.environ.pop(key, None)
    for key, val in cfg.items():
        os.environ[key] = str(val)
    sess.out.zero_()
    sess._launch(dom, [], complex(net.scalar), 0, stream)
    torch.cuda.synchronize()
    what = lib.b2_last_kernel().decode()
    same = bool(torch.equal(sess.out, ref))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sess._launch(dom, [], complex(net.scalar), 0, stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"cfg": cfg, "kernel": what, "ms": round(ms, 4),
                      "gbs": round(dom.bytes / ms / 1e6, 1), "bit_equal": same}), flush=True)
