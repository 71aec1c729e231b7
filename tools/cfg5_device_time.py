"""Device time of the batched cfg5a plan only (1024 parameter sets, one plan)."""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from optyx_b200 import workloads as cases  # noqa: E402
from optyx_b200 import executor  # noqa: E402
from optyx_b200.core import network as nw  # noqa: E402
from optyx_b200.core.backends import B200Backend  # noqa: E402

batch = 1024
backend = B200Backend()
make = lambda j: cases.fusion_link((math.cos(j * (math.pi / 2) / 1023),            # noqa: E731
                                    math.sin(j * (math.pi / 2) / 1023)), "bell")
nets, _ = backend.compile_batch(make, batch)
for net in nets:
    net.leaves.append(nw.Leaf(nw.K_DENSE, (), (), 0, np.array([net.scalar], dtype=complex), "s"))
    net.scalar = 1.0 + 0j
plan = executor.get_plan(nets[0], batch=batch)
sess = executor.Session(plan)
sess.set_leaves(nets)
sess.capture(1.0 + 0j)
for _ in range(3):
    sess.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    sess.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("cfg5a batch", batch, "steps", len(plan.steps), "device_ms_per_batch", round(ms, 4),
      "GB/s", round(plan.total_bytes / ms / 1e6, 1), flush=True)
