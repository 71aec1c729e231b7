"""Throughput of the GPU permanent backend: all output occupations
of n photons in m modes through a random interferometer."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optyx_b200 import photonic
from optyx_b200.core.backends import PermanentBackend

out = []
for m, n in ((8, 4), (12, 6), (12, 8), (16, 8), (14, 10)):
    d = photonic.Create(*([1] * n + [0] * (m - n))) >> photonic.seeded_ansatz(m, m, seed=0)
    pb = PermanentBackend()
    pb.prob_dist(d)                                   # warm-up (module load)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    p = pb.prob_dist(d)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    configs = len(list(p)) if False else None
    from math import comb
    nconf = comb(n + m - 1, n)
    flops = nconf * (2 ** (n - 1)) * (8.0 * n + 2)
    out.append({"modes": m, "photons": n, "configs": nconf, "total_prob": float(sum(p.values())),
                "e2e_ms": dt * 1e3, "kernel_ms": pb.last_kernel_ms,
                "kernel_tflops": flops / (pb.last_kernel_ms * 1e-3) / 1e12})
    print(json.dumps(out[-1]), flush=True)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from optyx_b200 import workloads as cases
for w, n, k in ((6, 3, 2), (8, 4, 4), (12, 6, 4)):
    d = cases.cfg3_interferometer(width=w, n_photons=n, measured=k, inflate=False)
    pb = PermanentBackend()
    pb.prob_dist_measured(d)
    t0 = time.perf_counter()
    p = pb.prob_dist_measured(d)
    dt = time.perf_counter() - t0
    flops = pb.last_configs * (2 ** (n - 1)) * (8.0 * n + 2)
    out.append({"workload": f"cfg3 lossy, distinguishable: {w} modes, {n} photons, {k} measured",
                "output_modes": 4 * w, "configs": pb.last_configs, "outcomes": len(p),
                "total_prob": float(sum(p.values())), "e2e_ms": dt * 1e3, "kernel_ms": pb.last_kernel_ms,
                "kernel_tflops": flops / (pb.last_kernel_ms * 1e-3) / 1e12})
    print(json.dumps(out[-1]), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_permanent.json"), "w"), indent=1)
