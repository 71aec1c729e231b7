"""Launch the dominant kernel of a bench workload a few times inside a
cudaProfilerStart/Stop window -- the command ncu captures for ``roofline.traffic``:

    python tools/prof_dominant.py cfg4 > gpurun_out/plain.log 2>&1 &&
    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -c 3 -o gpurun_out/r02_dom_cfg4 python tools/prof_dominant.py cfg4

Prints the kernel signature bench.py uses as the key of profiles/r02_traffic.json
(``tools/ncu_traffic.py`` turns the .ncu-rep into that table).  Extra arguments:
``leaf`` profiles the array build instead, ``small`` the small-step program.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from optyx_b200 import executor  # noqa: E402
from optyx_b200.core import network as nw  # noqa: E402
from optyx_b200.core.backends import B200Backend  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
what = sys.argv[2] if len(sys.argv) > 2 else "dom"
w = bench.workload(name)
kw = dict(target_size=w["target"], trials=w["trials"], min_slices=w.get("min_slices", 1))
backend = B200Backend(**kw)
if w["kind"] == "batch":
    nets, _ = backend.compile_batch(w["make"], w["batch"])
    for net in nets:
        net.leaves.append(nw.Leaf(nw.K_DENSE, (), (), 0, np.array([net.scalar], dtype=complex), "scalar"))
        net.scalar = 1.0 + 0j
    batch = w["batch"]
else:
    nets = [backend._get_network(w["factory"](), nested_eval=backend._nested_eval)]
    batch = 1
plan = executor.get_plan(nets[0], batch=batch, **kw)
sess = executor.Session(plan)
sess.set_leaves(nets)
sess.build_leaves()
sess.contract(nets[0].scalar, slice_range=(0, 1) if plan.sliced else None)
torch.cuda.synchronize()
stream = torch.cuda.current_stream().cuda_stream
hbm = bench.hbm_peak_gbs()
dom = max(plan.steps, key=lambda s: max(s.bytes / (hbm * 1e9), 8.0 * s.macs / (bench.FP64_PIPE_TFLOPS * 1e12)))
zeros = [0] * len(plan.sliced)
if what == "leaf":
    fn, sig = sess.build_leaves, "leaf_build"
elif what == "small":
    fn, sig = (lambda: sess._launch_small_groups(stream)), "small_program"
else:
    fn = lambda: sess._launch(dom, zeros, 1.0 + 0j, 0, stream)  # noqa: E731
    fn()
    sig = f"{sess.lib.b2_last_kernel().decode().split()[0]}:{'x'.join(str(x) for x in dom.mnk)}"
for _ in range(3):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
for _ in range(3):
    fn()
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
ms = e0.elapsed_time(e1) / 3
print(f"key={name}/{sig} kernel={sess.lib.b2_last_kernel().decode()} ms_per_launch={ms:.4f} "
      f"alg_bytes={dom.bytes if what == 'dom' else plan.arena_elems * 16 * batch} "
      f"macs={dom.macs if what == 'dom' else 0}")
