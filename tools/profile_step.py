"""Run the dominant (final) contraction step of the cfg2 bench workload a few
times with both kernels, for ncu.  usage: python tools/profile_step.py [k_open]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from optyx_b200 import workloads as cases  # noqa: E402
from optyx_b200 import executor  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 14
net = cases.compile_channel(cases.ghz(20, open_qubits=k, noise=0.95))
stream = torch.cuda.current_stream().cuda_stream
for thr in [(32, 16, 8, 4.0), (16, 16, 4)]:
    plan = executor.get_plan(net, gemm_threshold=thr)
    sess = executor.Session(plan)
    sess.run([net])
    dom = max(plan.steps, key=lambda s: s.bytes)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()          # ncu --profile-from-start off
    e0.record()
    for _ in range(3):
        sess._launch(dom, [], 1.0 + 0j, 0, stream)
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(dom.kernel, dom.mnk, "ms/launch", e0.elapsed_time(e1) / 3,
          "GB/s", dom.bytes / (e0.elapsed_time(e1) / 3 * 1e-3) / 1e9)
    del sess
    torch.cuda.empty_cache()
