"""Throughput of the other BASELINE configs (cfg1, cfg3 scaled, cfg5 batched) on one
GPU -> gpurun_out/configs.json.  bench.py stays the contract for the headline."""
import json
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from optyx_b200 import workloads as cases  # noqa: E402
from optyx_b200 import executor  # noqa: E402
from optyx_b200.core.backends import B200Backend  # noqa: E402

out = {}
results = {}
backend = B200Backend()


def _dump():
    os.makedirs("gpurun_out", exist_ok=True)
    name = "configs_only.json" if globals().get("ONLY") else "configs.json"
    json.dump(out, open(os.path.join("gpurun_out", name), "w"), indent=1)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def single(name, factory, reps=50, dtype=np.complex128, check_against=None):
    global backend
    backend = B200Backend(dtype=dtype)
    d = factory()
    t0 = time.perf_counter()
    net = backend._get_network(d, nested_eval=backend._nested_eval)
    t_compile = time.perf_counter() - t0
    t0 = time.perf_counter()
    plan = executor.get_plan(net, dtype=dtype)
    t_plan = time.perf_counter() - t0
    sess = executor.Session(plan)
    sess.set_leaves([net])
    sess.run([net])
    if len(plan.steps) > 8 and plan.parallel_safe:
        sess.capture(net.scalar)
        dt = timed(sess.replay, reps)
    else:
        dt = timed(lambda: (sess.build_leaves(), sess.contract(net.scalar)), reps)
    res = d.eval(backend)
    t0 = time.perf_counter()
    for _ in range(5):
        res = factory().eval(backend)
    t_e2e = (time.perf_counter() - t0) / 5
    t0 = time.perf_counter()
    p = res.prob_dist() if len(res.tensor.dom) == 0 and res.tensor.array.ndim > 0 else None
    t_prob = time.perf_counter() - t0
    out[name] = {"leaves": len(net.leaves), "steps": len(plan.steps), "out_shape": list(net.output_shape),
                 "macs": plan.total_macs, "bytes": plan.total_bytes,
                 "device_ms_per_eval": dt * 1e3, "device_evals_per_s": 1 / dt,
                 "e2e_ms_per_eval": t_e2e * 1e3, "host_compile_ms": t_compile * 1e3,
                 "host_plan_ms": t_plan * 1e3, "prob_dist_ms": t_prob * 1e3,
                 "prob_total": float(sum(p.values())) if p else None,
                 "tflops": 8 * plan.total_macs / dt / 1e12, "gbs": plan.total_bytes / dt / 1e9,
                 "dtype": np.dtype(dtype).name,
                 "kernels": {k: sum(1 for st in plan.steps if st.kernel == k) for k in ("direct", "gemm")}}
    results[name] = res.tensor.array
    if check_against is not None:
        ref = results[check_against]
        out[name]["max_rel_err_vs_" + check_against] = float(
            np.abs(res.tensor.array - ref).max() / max(1e-300, np.abs(ref).max()))
    print(name, out[name], flush=True)
    _dump()


def batched(name, make, batch=1024, reps=10):
    t0 = time.perf_counter()
    ds = [make(j) for j in range(batch)]
    nets = [backend._get_network(d, nested_eval=backend._nested_eval) for d in ds]
    t_compile = time.perf_counter() - t0
    from optyx_b200.core import network as nw
    for net in nets:
        net.leaves.append(nw.Leaf(nw.K_DENSE, (), (), 0, np.array([net.scalar], dtype=complex), "s"))
        net.scalar = 1.0 + 0j
    plan = executor.get_plan(nets[0], batch=batch)
    sess = executor.Session(plan)
    t0 = time.perf_counter()
    sess.set_leaves(nets)
    torch.cuda.synchronize()
    t_upload = time.perf_counter() - t0
    sess.capture(1.0 + 0j)
    dt = timed(sess.replay, reps)
    arena_bytes = plan.arena_elems * 16 * batch
    t_leaf = timed(sess.build_leaves, reps)
    t0 = time.perf_counter()
    res = backend.eval_batch(ds)
    t_e2e_serial = time.perf_counter() - t0
    backend.eval_batch(make=make, count=batch)               # warm (plan cache, pinned buffers)
    t0 = time.perf_counter()
    res = backend.eval_batch(make=make, count=batch)         # host compile on a fork pool
    t_e2e = time.perf_counter() - t0
    out[name] = {"batch": batch, "leaves": len(nets[0].leaves), "steps": len(plan.steps),
                 "device_ms_per_batch": dt * 1e3, "device_evals_per_s": batch / dt,
                 "leaf_build_ms": t_leaf * 1e3, "leaf_build_gbs": arena_bytes / t_leaf / 1e9,
                 "arena_bytes": arena_bytes, "plan_bytes": plan.total_bytes, "gbs": plan.total_bytes / dt / 1e9,
                 "host_compile_s": t_compile, "params_upload_ms": t_upload * 1e3,
                 "e2e_evals_per_s": batch / t_e2e, "e2e_evals_per_s_serial_compile": batch / t_e2e_serial,
                 "host_cores": os.cpu_count(), "first": complex(res[0].tensor.array.reshape(-1)[0]).real}
    print(name, out[name], flush=True)
    _dump()


ONLY = sys.argv[1] if len(sys.argv) > 1 else ""
if ONLY:                                   # e.g. `bench_configs.py cfg5`: only matching rows
    _single, _batched = single, batched
    single = lambda name, *a, **k: _single(name, *a, **k) if ONLY in name else None       # noqa: E731
    batched = lambda name, *a, **k: _batched(name, *a, **k) if ONLY in name else None     # noqa: E731
single("cfg1_hom_lossy", lambda: cases.lossy_hom(0.8))
single("cfg1_6mode_3photon", lambda: cases.cfg1_interferometer(6, (1, 1, 1, 0, 0, 0)))
single("cfg3_6mode_3photon_k2", lambda: cases.cfg3_interferometer(width=6, n_photons=3, measured=2), reps=3)
single("cfg3_6mode_3photon_k2_complex64", lambda: cases.cfg3_interferometer(width=6, n_photons=3, measured=2),
       reps=3, dtype=np.complex64, check_against="cfg3_6mode_3photon_k2")
single("cfg2_ghz20_measured", lambda: cases.ghz(20, noise=0.95, measure=True))
backend = B200Backend()
batched("cfg5a_fusion_link_x1024",
        lambda j: cases.fusion_link((math.cos(j * (math.pi / 2) / 1023), math.sin(j * (math.pi / 2) / 1023)), "bell"))
batched("cfg5b_fusion_teleport_x1024", lambda j: cases.fusion_teleport_input(j / 1024))
def permanents(name, w, n, k, reps=5):
    """cfg3 through PermanentBackend.prob_dist_measured (loss + partial distinguishability)."""
    from math import comb
    from optyx_b200.core.backends import PermanentBackend
    d = cases.cfg3_interferometer(width=w, n_photons=n, measured=k, inflate=False)
    pb = PermanentBackend()
    pb.prob_dist_measured(d)
    t0 = time.perf_counter()
    for _ in range(reps):
        p = pb.prob_dist_measured(cases.cfg3_interferometer(width=w, n_photons=n, measured=k, inflate=False))
    t_e2e = (time.perf_counter() - t0) / reps
    flops = pb.last_configs * (2 ** (n - 1)) * (8.0 * n + 2)
    out[name] = {"modes": w, "photons": n, "measured": k, "output_modes": 4 * w,
                 "configurations": pb.last_configs, "outcomes": len(p),
                 "prob_total": float(sum(p.values())), "device_ms_per_eval": pb.last_kernel_ms,
                 "e2e_ms_per_eval": t_e2e * 1e3,
                 "fp64_tflops": flops / (pb.last_kernel_ms * 1e-3) / 1e12}
    print(name, out[name], flush=True)
    _dump()


if not ONLY or ONLY in "cfg3_full_12mode_6photon_k4_permanents":
    permanents("cfg3_full_12mode_6photon_k4_permanents", 12, 6, 4)
_dump()
