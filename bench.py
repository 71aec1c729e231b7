#!/usr/bin/env python
"""Benchmark of the hot path: ``diagram.eval()`` diagrams/sec on the BASELINE.json
configs, with the roofline of the dominant kernel and the CPU baseline beside it.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

A "step" is one pass of the hot path over one diagram: leaf-array build + every
pairwise contraction of the plan (device-resident descriptors; ``value``), and,
for ``e2e``, the whole user call ``diagram.eval(B200Backend(...))`` from host
objects to a host numpy result (host compile + H2D of descriptors/params + kernels +
D2H of the result inside the timed region).

Headline workload (default, every N): **cfg4**, the 30-qubit Clifford+T amplitude
network whose cotengra-style index slices are the units that shard over GPUs -- the one
BASELINE config with a real exchange step (ONE NCCL sum): the same line measures 1, 2,
4 and 8 GPUs, total work fixed (``"scaling": "strong"``), so the driver's efficiency
is the scaling of north_star piece (3).  At N = 1 the line also carries the complete
measurement of cfg2 (the 2^28-entry density tensor, HBM-bound) under ``"cfg2"``, and at
N > 1 its replica throughput under ``"replicas"``.  ``--workload cfg1|cfg2|cfg3|cfg5a|
cfg5b`` prints the same schema for the other configs (profiles/r02_bench_*.json).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC, UNIT = "eval_diagrams_per_sec", "diagrams/s"
FP64_PIPE_TFLOPS = 37.0      # DMMA m8n8k4 / DFMA, measured on this pool (profiles/r01_probe.json)
THREE_M_VARIANTS = ("gemm_ws", "gemm_ws32", "gemm_ws16", "gemm_ws_splitk", "gemm3m", "gemm3m_splitk",
                    "gemm_thin")      # kernels that multiply complex numbers with 3 real products
CFG4_TARGET = float(2 ** 27)       # <= 2^27 elements (2 GiB complex128) per slice
CFG4_TRIALS = 128                  # path-search budget (cotengra's default hyper-optimiser: 128 repeats)
CFG4_MIN_SLICES = 32               # ... and >= 32 slices, so that 8 GPUs get >= 4 each


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        return {}


def hbm_peak_gbs():
    return float(peaks().get("hbm_gbs", 6650.0))        # fallback: B200_PROFILING.md


def workload(name: str):
    """name -> dict(factory, desc, kind, ...).  kinds: "sliced" (slices shard over the
    ranks), "single" (one diagram per step), "batch" (1024 structurally identical
    diagrams per step)."""
    from optyx_b200 import workloads as cases
    if name == "cfg4":
        n, depth = 30, 48
        return dict(kind="sliced", factory=lambda: cases.random_clifford_t(n, depth, seed=0),
                    target=CFG4_TARGET, trials=CFG4_TRIALS, min_slices=CFG4_MIN_SLICES,
                    desc=f"cfg4: {n}-qubit brickwork Clifford+T depth {depth} amplitude <0|C|0>, "
                         f"sliced to <= 2^{int(math.log2(CFG4_TARGET))} elements per slice, "
                         f">= {CFG4_MIN_SLICES} slices")
    if name == "cfg2":
        n, k = 20, 14
        return dict(kind="single", factory=lambda: cases.ghz(n, open_qubits=k, noise=0.95),
                    target=None, trials=8,
                    desc=f"cfg2: {n}-qubit GHZ, DephasingError(0.95) on every qubit, CP-doubled, "
                         f"{k} qubits open / {n - k} discarded -> (2,)^{2 * k} density tensor")
    if name == "cfg1":
        return dict(kind="single", factory=lambda: cases.cfg1_interferometer(6, (1, 1, 1, 0, 0, 0)),
                    target=None, trials=8,
                    desc="cfg1: 6-mode 3-photon lossy interferometer, PhotonLoss(0.9), ansatz(6,6) "
                         "rng(0), NumberResolvingMeasurement(6), truncation dim 4")
    if name == "cfg3":
        return dict(kind="single", target=None, trials=8,
                    factory=lambda: cases.cfg3_interferometer(width=6, n_photons=3, measured=2),
                    desc="cfg3 (tensor-network instance): 6-mode 3-photon lossy interferometer, "
                         "partially distinguishable photons, inflate(2), 2 modes measured / 4 discarded")
    if name in ("cfg5a", "cfg5b"):
        batch = 1024
        if name == "cfg5a":
            def make(j):
                th = j * (math.pi / 2) / (batch - 1)
                return cases.fusion_link((math.cos(th), math.sin(th)), "bell")
            desc = "cfg5a: heralded fusion link (README.md:230-280), inflate(2), 1024 internal-state angles"
            params = (np.arange(batch) * (math.pi / 2) / (batch - 1),)

            def make_arr(th):                 # ONE diagram with array-valued amplitudes
                return cases.fusion_link((np.cos(th), np.sin(th)), "bell")
        else:
            def make(j):
                return cases.fusion_teleport_input(j / batch)
            desc = "cfg5b: fusion-II teleportation with classical feed-forward, 1024 input phases"
            params = (np.arange(batch) / batch,)
            make_arr = cases.fusion_teleport_input
        return dict(kind="batch", make=make, make_arr=make_arr, params=params, batch=batch,
                    target=None, trials=8, desc=desc)
    raise SystemExit(f"unknown workload {name}")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "25"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t0 = time.time()                 # nvidia-smi takes a moment to print its first line
            while not self.lines and time.time() - t0 < 3.0:
                time.sleep(0.01)
            self.lines.clear()               # keep only samples taken under load
        except OSError:
            self.proc = None
        return self

    def hold_load(self, fn, sync, seconds=0.3, min_samples=4):
        """Keep the same kernels running (untimed) until enough samples exist: the
        timed region itself can be shorter than one sampling period."""
        t0 = time.time()
        while self.proc is not None and (len(self.lines) < min_samples or time.time() - t0 < seconds):
            fn()
            sync()
            if time.time() - t0 > 3.0:
                break

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = max(mx, float(parts[1]))
            except ValueError:
                continue
            for nme, flag in zip(names, parts[2:6]):
                if flag.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------
# CPU arm: the reference's path on the host cores
# --------------------------------------------------------------------------------------
def reference_probe():
    """is the real reference importable on this box?"""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref_dir) and ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        import cotengra  # noqa: F401
        import discopy  # noqa: F401
        import optyx  # noqa: F401
        import quimb  # noqa: F401
        return True, ""
    except Exception as exc:
        return False, f"{type(exc).__name__}: {exc}"


def cpu_reference(name: str, budget_s: float = 25.0, max_evals: int = 3):
    """The reference's CPU path for this metric on a BOUNDED sample of the workload.

    The reference itself (QuimbBackend + cotengra) cannot be installed (discopy, quimb,
    cotengra are absent from this image, no network: ``reference_probe``), so this is
    the oracle PORT: ``oracle/compile.py`` (CP doubling, light-cone truncation, dense
    arrays -- the reference's host compile restated) + ``oracle/contract.py`` (pairwise
    transpose + BLAS zgemm, all host threads).  cfg4: the same sliced tree the GPU
    executes, ONE slice timed and scaled by the slice count (each slice costs the same).
    Returns (cpu_baseline dict, seconds per eval, evals actually timed)."""
    from oracle import compile as oc
    from oracle.contract import contract_along_path, contract_arrays
    w = workload(name)
    cores = os.cpu_count()
    # torchrun exports OMP_NUM_THREADS=1 to every rank: give the CPU arm all host threads back,
    # so that the baseline is the same at every N (r01: 6.0 -> 7.4 s per eval under torchrun)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:
        pass
    if w["kind"] == "sliced":
        from optyx_b200 import executor
        from optyx_b200.core.backends import B200Backend
        d = w["factory"]()
        t0 = time.perf_counter()
        tn, out, _, _ = oc.compile_dense(d)                       # the reference-shaped host compile
        t_compile = time.perf_counter() - t0
        # the tree: found on the product's network (hyper-indices) -- the numpy port
        # contracts that same network so that it does the same arithmetic per slice
        net = B200Backend()._get_network(d)
        plan = executor.get_plan(net, target_size=w["target"], trials=w["trials"],
                                 min_slices=w.get("min_slices", 1))
        from oracle import leaves as ol
        arrays = ol.build_all(net.leaves)
        inds = [list(l.inds) for l in net.leaves]
        t0 = time.perf_counter()
        contract_along_path(arrays, inds, net.dims, net.output_inds, plan.info.path,
                            {q: 0 for q in plan.sliced}, net.scalar)
        t_slice = time.perf_counter() - t0
        secs = t_compile + t_slice * plan.nslices
        sample = (f"{w['desc']}: host compile ({t_compile:.2f} s) + 1 of {plan.nslices} slices "
                  f"({t_slice:.2f} s, slice-invariant part included) x {plan.nslices}, numpy/BLAS "
                  f"threads = all cores")
        return ({"value": 1.0 / secs, "unit": UNIT, "cores": cores, "kind": "port",
                 "sample": sample}, secs, 1)
    if w["kind"] == "batch":
        times = []
        for j in range(min(max_evals * 4, w["batch"])):
            t0 = time.perf_counter()
            oc.eval_dense(w["make"](j))
            times.append(time.perf_counter() - t0)
            if sum(times) > budget_s:
                break
        secs = float(np.mean(times))
        return ({"value": 1.0 / secs, "unit": UNIT, "cores": cores, "kind": "port",
                 "sample": f"{w['desc']}: {len(times)} of {w['batch']} evals, host compile + "
                           "contraction, numpy/BLAS threads = all cores"}, secs, len(times))
    times = []
    for _ in range(max_evals):
        t0 = time.perf_counter()
        tn, out, _, _ = oc.compile_dense(w["factory"]())
        contract_arrays(tn.arrays, tn.inds, tn.dims, out, tn.scalar)
        times.append(time.perf_counter() - t0)
        if sum(times) > budget_s:
            break
    secs = float(np.mean(times))
    return ({"value": 1.0 / secs, "unit": UNIT, "cores": cores, "kind": "port",
             "sample": f"{w['desc']} ({len(times)} evals, host compile + contraction, numpy/BLAS "
                       "threads = all cores)"}, secs, len(times))


def reference_arm(args):
    ok, why = reference_probe()
    base, secs, n_timed = cpu_reference(args.workload)
    w = workload(args.workload)
    base["reference_importable"] = ok
    if not ok:
        base["reference_import_error"] = why[:160]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": n_timed, "warmup": 0, "steps_requested": args.steps,
        "ms_per_step": secs * 1e3, "higher_is_better": True,
        "scaling": "strong" if w["kind"] == "sliced" else "weak",
        "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": {"workload": w["desc"], "l2": "inputs larger than L2 / not applicable on the CPU"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0}}))


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        from optyx_b200 import runtime as rt
        self.lib = rt.load()
        self.hbm = hbm_peak_gbs()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, secs: float) -> float:
        if self.world == 1:
            return secs
        tt = self.torch.tensor([secs], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(tt, op=self.dist.ReduceOp.MAX)
        return float(tt.item())

    def timed(self, fn, reps: int) -> float:
        """Seconds per call: CUDA events on the launching stream, sync on both sides."""
        torch = self.torch
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / reps

    # ---- one workload -------------------------------------------------------------------
    def measure(self, name: str, steps: int, warmup: int, sharded: bool, with_cpu: bool):
        torch = self.torch
        from optyx_b200 import dist as b2dist
        from optyx_b200 import executor
        from optyx_b200.core import network as nw
        from optyx_b200.core.backends import B200Backend
        from optyx_b200.planner import plan_bound_seconds
        w = workload(name)
        world = self.world if sharded else 1
        backend_kw = dict(target_size=w["target"], trials=w["trials"],
                          min_slices=w.get("min_slices", 1))
        compile_backend = B200Backend(**backend_kw)
        t0 = time.perf_counter()
        if w["kind"] == "batch":
            batch = w["batch"]
            nets, _ = compile_backend.compile_batch(w["make"], batch)
            for net in nets:                  # per-eval scalars travel as a rank-0 leaf
                net.leaves.append(nw.Leaf(nw.K_DENSE, (), (), 0,
                                          np.array([net.scalar], dtype=np.complex128), "scalar"))
                net.scalar = 1.0 + 0j
            net, diagram = nets[0], None
        else:
            batch = 1
            diagram = w["factory"]()
            net = compile_backend._get_network(diagram, nested_eval=compile_backend._nested_eval)
            nets = [net]
        t_compile = time.perf_counter() - t0
        t0 = time.perf_counter()
        plan = executor.get_plan(net, batch=batch, **backend_kw)
        t_plan = time.perf_counter() - t0
        sess = executor.Session(plan)
        sess.set_leaves(nets)
        torch.cuda.synchronize()
        scalar = net.scalar

        if plan.sliced:
            def step_fn():
                if world > 1:
                    b2dist.run_sliced(sess, net, upload=False)
                else:
                    sess.build_leaves()
                    sess.contract(scalar)
            use_graph = False
        else:
            use_graph = len(plan.steps) > 8
            if use_graph:
                sess.capture(scalar)
                step_fn = sess.replay
            else:
                def step_fn():
                    sess.build_leaves()
                    sess.contract(scalar)
        sess.launches = 0
        sess.build_leaves()
        if plan.sliced and world > 1:
            sess.contract(scalar, slice_range=b2dist.partition(plan.nslices, self.rank, world))
        else:
            sess.contract(scalar)
        launches_per_step = sess.launches
        torch.cuda.synchronize()

        for _ in range(warmup):
            step_fn()
        self.barrier()
        with ClockSampler(self.local_rank) as clocks:
            step_fn()                            # the sampler start-up left the GPU idle
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            profiled = os.environ.get("B2_PROFILE_TIMED") == name    # ncu --profile-from-start off
            if profiled:
                torch.cuda.profiler.start()
            e0.record()
            for _ in range(steps):
                step_fn()
            e1.record()
            if profiled:
                torch.cuda.synchronize()
                torch.cuda.profiler.stop()
            self.barrier()
            secs = self.max_over_ranks(e0.elapsed_time(e1) * 1e-3)
            # ---- dominant kernel alone, same stream, CUDA events over K launches
            dom = max(plan.steps, key=lambda s: max(s.bytes / (self.hbm * 1e9),
                                                    8.0 * s.macs / (FP64_PIPE_TFLOPS * 1e12)))
            zeros = [0] * len(plan.sliced)
            stream = torch.cuda.current_stream().cuda_stream
            for _ in range(3):
                sess._launch(dom, zeros, 1.0 + 0j, 0, stream)
            dom_secs = self.timed(lambda: sess._launch(dom, zeros, 1.0 + 0j, 0, stream), max(steps, 5))
            dom_variant = self.lib.b2_last_kernel().decode().split()[0]
            leaf_secs = self.timed(sess.build_leaves, max(steps, 20))
            clocks.hold_load(step_fn, torch.cuda.synchronize)
        per_step = secs / steps
        value = batch * (world if not plan.sliced else 1) / per_step

        # ---- result check inside the bench (cheap, size-independent): slices sum to the
        #      same amplitude on every N; the batch / single results are finite
        out_host = None
        if plan.out_elems <= (1 << 16):
            out_host = sess.out.reshape(-1)[:min(4, plan.out_elems * batch)].cpu().numpy()

        # ---- e2e: the user call, host objects in, host numpy out
        e2e = self.time_e2e(w, backend_kw, sharded, steps)

        tensor_bound = 8.0 * dom.macs / (FP64_PIPE_TFLOPS * 1e12) > dom.bytes / (self.hbm * 1e9)
        if tensor_bound:
            achieved, peak_val, unit = 8.0 * dom.macs / dom_secs / 1e12, FP64_PIPE_TFLOPS, "TFLOP/s"
        else:
            achieved, peak_val, unit = dom.bytes / dom_secs / 1e9, self.hbm, "GB/s"
        kernel_sig = f"{dom_variant}:{'x'.join(str(x) for x in dom.mnk)}"
        three_m = dom_variant in THREE_M_VARIANTS
        roofline = {
            "bound": "tensor" if tensor_bound else "hbm", "achieved": achieved, "peak": peak_val,
            "unit": unit, "frac": achieved / peak_val,
            "traffic": measured_traffic(name, kernel_sig),
            "peak_source": ("FP64 DMMA peak measured on this pool (profiles/r01_probe.json)"
                            if tensor_bound else
                            "MEASURED_PEAKS.json hbm_gbs" if peaks() else "fallback 6650 GB/s"),
            "kernel": f"b2_contract_{dom.kernel} [{dom_variant}] ({'final step, ' if dom.final else ''}"
                      f"(batch,M,N,K)={dom.mnk})",
            "kernel_signature": kernel_sig,
            "algorithmic_bytes_per_launch": dom.bytes, "launch_ms": dom_secs * 1e3,
            "macs_per_launch": dom.macs,
            "share_of_step": dom_secs * (plan.nslices / world if dom.per_slice else 1) / per_step,
        }
        if tensor_bound and three_m:
            # 3M complex multiplication: the kernel EXECUTES 6 flops per complex multiply-add
            # on the FP64 pipe for the 8 algorithmic ones, so `frac` (algorithmic / pipe peak)
            # can exceed 1; pipe_frac is what the tensor pipe itself is doing
            roofline["executed_flops_per_launch"] = 6.0 * dom.macs
            roofline["pipe_frac"] = 6.0 * dom.macs / dom_secs / 1e12 / FP64_PIPE_TFLOPS
            roofline["note"] = ("3M complex multiplication (3 real DMMA products per complex product): "
                                "achieved/frac count the 8 algorithmic flops per complex multiply-add "
                                "(SURVEY.md 8(d)), pipe_frac the 6 executed ones")
        bound_s = plan_bound_seconds(plan, self.hbm, FP64_PIPE_TFLOPS) / (world if plan.sliced else 1)
        arena_bytes = plan.arena_elems * plan.dtype.itemsize * batch
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
            "scaling": "strong" if plan.sliced else "weak", "vs_baseline": None, "dtype": "c128",
            "data": "synthetic",
            "config": {"workload": w["desc"], "leaves": len(net.leaves), "plan_steps": len(plan.steps),
                       "batch": batch, "nslices": plan.nslices, "peak_elems": plan.info.peak,
                       "cuda_graph": bool(use_graph),
                       "l2": (f"largest intermediate {plan.info.peak * 16 / 1e6:.0f} MB, output "
                              f"{plan.out_elems * 16 * batch / 1e6:.0f} MB per step: working set "
                              "larger than the 126 MB L2") if plan.info.peak * 16 > 126e6
                       else "working set fits L2: launch-bound config, no roofline claim per eval",
                       "plan_macs": plan.total_macs, "plan_bytes": plan.total_bytes,
                       "host_compile_ms": t_compile * 1e3, "host_plan_ms": t_plan * 1e3},
            "clocks": clocks.summary(),
            "e2e": e2e,
            "gpu_launches": launches_per_step * steps,
            "roofline": roofline,
            "plan_roofline": {"bound_ms": bound_s * 1e3, "measured_ms": per_step * 1e3,
                              "frac": bound_s / per_step, "hbm_gbs": self.hbm,
                              "fp64_pipe_tflops": FP64_PIPE_TFLOPS,
                              "note": "sum over plan steps of max(algorithmic bytes / HBM peak, "
                                      "8 MACs / FP64 DMMA peak) / measured step time"},
            "contraction_tflops": 8.0 * plan.total_macs * (1 if plan.sliced else world)
            / per_step / 1e12,
            "leaf_build": {"ms": leaf_secs * 1e3, "bytes": arena_bytes,
                           "gbs": arena_bytes / leaf_secs / 1e9,
                           "frac_of_hbm": arena_bytes / leaf_secs / 1e9 / self.hbm,
                           "note": ("one launch builds every leaf array of the step; below a few MB it "
                                    "is launch-bound (~10 us) and the GB/s figure says nothing -- the "
                                    "bandwidth of the array build is on the cfg5a line (58 MB batched "
                                    "arena, profiles/r02_bench_cfg5a.json)")},
        }
        if out_host is not None:
            line["result_head"] = [[float(v.real), float(v.imag)] for v in out_host]
        if with_cpu and self.rank == 0:
            line["cpu_baseline"], _, _ = cpu_reference(name)
        del sess
        torch.cuda.empty_cache()
        return line

    def time_e2e(self, w, backend_kw, sharded: bool, steps: int):
        torch = self.torch
        from optyx_b200.core.backends import B200Backend
        world = self.world if sharded else 1
        backend = B200Backend(distributed=(sharded and self.world > 1), **backend_kw)
        if w["kind"] == "batch":
            def call():                      # one compile with array-valued parameters
                return backend.eval_params(w["make_arr"], *w["params"])
            units = w["batch"]
        else:
            fresh = []                           # one NEW diagram object per timed call, built before

            def call():
                d = fresh.pop() if fresh else w["factory"]()
                return d.eval(backend)
            units = 1
        res = None
        for _ in range(3):                       # warm-up: plan cache + pinned host buffers
            res = call()
            _touch(res)
        n_e2e = max(2, min(steps, 5))
        if w["kind"] != "batch":
            # the timed call is diagram.eval(backend) -- the call a user makes -- on objects the
            # backend has never seen (nothing cached on them); constructing the circuit (user
            # code, 15 ms of Python for cfg4's 3588 boxes) is not part of eval()
            fresh.extend(w["factory"]() for _ in range(n_e2e))
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            res = call()
            _touch(res)                          # the host array is materialised inside the timed region
        torch.cuda.synchronize()
        secs = self.max_over_ranks((time.perf_counter() - t0) / n_e2e)
        info = dict(B200Backend.last_d2h)
        h2d = int(getattr(backend, "last_h2d_bytes", 0))
        sliced = w["kind"] == "sliced"
        return {"value": units * (1 if sliced else world) / secs, "unit": UNIT,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(info.get("bytes", 0)),
                "ms_per_eval": secs * 1e3 / units, "d2h": info.get("mode"),
                "note": "host diagram -> host numpy array through diagram.eval(B200Backend(...)) on a "
                        "diagram object built just before and never seen by the backend (eval_params "
                        "for cfg5): structural fingerprint / host compile, H2D of descriptors + "
                        "host-computed leaf values, kernels, D2H of the result tensor"}

    # ---- piece (4): prob kernels against HBM ----------------------------------------------
    def measure_probs(self):
        """b2_probs_amp / b2_probs_dm on a 2^28-entry tensor (4.3 GB, the cfg2 result size):
        algorithmic bytes s*|T| + 8*|P| over the measured launch time."""
        import ctypes as C
        torch = self.torch
        from optyx_b200 import runtime as rt
        n = 1 << 28
        t = torch.empty(n, dtype=torch.complex128, device="cuda")
        t.real.uniform_(-1, 1)
        t.imag.zero_()
        p = torch.empty(n, dtype=torch.float64, device="cuda")
        nz = torch.empty(n, dtype=torch.uint8, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        amp = self.timed(lambda: rt.check(self.lib.b2_probs_amp(
            n, t.data_ptr(), p.data_ptr(), nz.data_ptr(), 0, stream)), 5)
        # DM layout: 4 measured bits + 12 unmeasured qubits (ket, bra pairs) = 28 axes of 2
        kinds = [1] * 4 + [2, 3] * 12
        nd = len(kinds)
        pm = torch.empty(16, dtype=torch.float64, device="cuda")
        seen = torch.empty(16, dtype=torch.uint8, device="cuda")
        mx = torch.empty(1, dtype=torch.float64, device="cuda")
        dims = (C.c_int32 * nd)(*([2] * nd))
        kk = (C.c_int32 * nd)(*kinds)
        dm = self.timed(lambda: rt.check(self.lib.b2_probs_dm(
            nd, dims, kk, t.data_ptr(), pm.data_ptr(), seen.data_ptr(), mx.data_ptr(), 0,
            stream)), 5)
        out = {"tensor_bytes": n * 16,
               "amp": {"ms": amp * 1e3, "gbs": (n * 16 + n * 8) / amp / 1e9},
               "dm": {"ms": dm * 1e3, "gbs": (n * 16 + 16 * 8) / dm / 1e9,
                      "layout": "4 measured bits + 12 unmeasured qubits"}}
        for k in ("amp", "dm"):
            out[k]["frac_of_hbm"] = out[k]["gbs"] / self.hbm
        del t, p, nz
        torch.cuda.empty_cache()
        return out


def _touch(res):
    """Force the host result of an eval (lazy results materialise on first access)."""
    if isinstance(res, list):
        return [np.asarray(r.tensor.array).reshape(-1)[:1] for r in res][-1]
    return np.asarray(res.tensor.array).reshape(-1)[:1]


def measured_traffic(name: str, kernel_sig: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch,
    from the committed ncu capture of THIS workload (profiles/r02_traffic.json, written
    by tools/ncu_traffic.py from the ncu csv); null when no capture of this kernel exists."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except (OSError, ValueError):
        return None
    hit = table.get(f"{name}/{kernel_sig}")
    return float(hit["dram_bytes"]) if hit else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="cfg4")
    ap.add_argument("--no-also", action="store_true", help="headline workload only")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args)
        return

    # rank 0 must print exactly ONE line on stdout: libraries (NCCL's version banner and
    # its NCCL_DEBUG log, ...) write to the C-level fd 1, so fd 1 is pointed at stderr
    # for the whole run and the JSON line goes to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    b = Bench(args)
    headline = workload(args.workload)
    sharded = headline["kind"] == "sliced"
    out = b.measure(args.workload, args.steps, args.warmup, sharded=sharded,
                    with_cpu=(b.world == 1 and not args.no_cpu))
    if not args.no_also and args.workload == "cfg4":
        k = max(3, min(args.steps, 10))
        if b.world == 1:
            out["cfg2"] = b.measure("cfg2", k, 3, sharded=False, with_cpu=not args.no_cpu)
            out["probs"] = b.measure_probs()
        else:
            rep = b.measure("cfg2", k, 3, sharded=False, with_cpu=False)
            # every rank evaluates its own diagrams: replicas, no collective
            vals = b.torch.tensor([rep["value"], rep["e2e"]["value"]], device="cuda",
                                  dtype=b.torch.float64)
            b.dist.all_reduce(vals)
            out["replicas"] = {"workload": rep["config"]["workload"], "scaling": "weak",
                               "value": float(vals[0]), "e2e_value": float(vals[1]),
                               "unit": UNIT, "ms_per_step": rep["ms_per_step"],
                               "note": "independent evals, one per rank per step, no collective"}
    if b.world > 1:
        b.dist.barrier()
        b.torch.cuda.synchronize()
        b.dist.destroy_process_group()
    sys.stdout.flush()
    if rank == 0:
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    os.close(real_stdout)


if __name__ == "__main__":
    main()
