"""How sparse are the result tensors?  (exact zeros vs tiny residues)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from optyx_b200 import workloads as cases  # noqa: E402
from optyx_b200 import executor  # noqa: E402

for name, d in (("ghz20_open11_noise", cases.ghz(20, open_qubits=11, noise=0.95)),
                ("ghz20_open11", cases.ghz(20, open_qubits=11)),
                ("ghz20_measured_noise", cases.ghz(20, noise=0.95, measure=True)),
                ("cfg1_6mode", cases.cfg1_interferometer(6, (1, 1, 1, 0, 0, 0)))):
    dev = executor.evaluate(cases.compile_channel(d)).reshape(-1)
    mag = dev.abs()
    mx = float(mag.max())
    print(name, "n", dev.numel(), "exact nonzero", int((mag != 0).sum()),
          "> 1e-12 max", int((mag > 1e-12 * mx).sum()), "max", mx,
          "largest residue", float(mag[mag <= 1e-12 * mx].max()) if (mag <= 1e-12 * mx).any() else None,
          flush=True)
