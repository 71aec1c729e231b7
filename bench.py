#!/usr/bin/env python
"""Benchmark of the hot path: ``diagram.eval()`` diagrams/sec on the BASELINE.json
configs, with the roofline of the dominant kernel and the CPU baseline beside it.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg2]

A "step" is one pass of the hot path over one diagram: leaf-array build +
every pairwise contraction of the plan (device-resident descriptors; `value`),
and, for `e2e`, the whole user call ``diagram.eval(B200Backend())`` from host
objects to a host numpy result (compile + H2D of descriptors/params + kernels +
D2H of the result tensor inside the timed region).

N > 1 (torchrun): evals are independent units -> every rank evaluates its own
diagrams, no data-path collective, ``"scaling": "weak"``.  The sliced network
(cfg4) is additionally run with its slices partitioned over the ranks and one
NCCL all-reduce; its time is reported under ``"sliced"`` at every N.
"""
from __future__ import annotations

import argparse
import json
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


def workload(name: str, small: bool = False):
    """(diagram factory, description).  ``small``: the bounded CPU sample."""
    from optyx_b200 import workloads as cases
    if name == "cfg2":
        n, k = (20, 14)        # the CPU arm runs the full-size workload too (~20 s/eval)
        return (lambda: cases.ghz(n, open_qubits=k, noise=0.95),
                f"cfg2: {n}-qubit GHZ, DephasingError(0.95) on every qubit, CP-doubled, "
                f"{k} qubits open / {n - k} discarded -> (2,)^{2 * k} density tensor")
    if name == "cfg1":
        return (lambda: cases.cfg1_interferometer(6, (1, 1, 1, 0, 0, 0)),
                "cfg1: 6-mode 3-photon lossy interferometer, PhotonLoss(0.9), ansatz(6,6) "
                "rng(0), NumberResolvingMeasurement(6), truncation dim 4")
    if name == "cfg4":
        n, depth = (30, 41) if not small else (20, 12)
        return (lambda: cases.random_clifford_t(n, depth, seed=0),
                f"cfg4: {n}-qubit brickwork Clifford+T depth {depth} amplitude <0|C|0>")
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
            for _ in range(20):
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


def cpu_reference(name: str, steps: int, warmup: int):
    """The reference's CPU path for this metric: numpy tensordot/matmul (BLAS,
    all host threads) along a greedy order -- the oracle port of what
    QuimbBackend does (the reference itself cannot be installed: discopy, quimb,
    cotengra are absent from this image and there is no network)."""
    from cases import compile_channel
    from oracle.contract import contract_network
    factory, desc = workload(name, small=True)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        net = compile_channel(factory())
        contract_network(net)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        if sum(times) > 25.0 and len(times) >= 1:
            break
    return {"value": 1.0 / float(np.mean(times)), "unit": UNIT, "cores": os.cpu_count(),
            "kind": "port", "sample": desc + f" ({len(times)} evals, compile + contraction, "
            "numpy/BLAS threads = all cores)"}, float(np.mean(times))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--no-sliced", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        base, t = cpu_reference(args.workload, max(1, min(args.steps, 3)), 0)
        _, desc = workload(args.workload, small=True)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": {"workload": desc, "note": "bounded CPU sample of the bench workload"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    from optyx_b200 import executor, runtime as rt
    from optyx_b200.core.backends import B200Backend

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    # rank 0 must print exactly ONE line on stdout: libraries (NCCL's version
    # banner, ...) write to the C-level fd 1, so fd 1 is pointed at stderr for the
    # whole run and the JSON line goes to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        if "B2_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["B2_NCCL_DEBUG"]
        else:
            os.environ.pop("NCCL_DEBUG", None)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = rt.load()

    factory, desc = workload(args.workload)
    diagram = factory()
    t0 = time.perf_counter()
    compile_backend = B200Backend()
    net = compile_backend._get_network(diagram, nested_eval=compile_backend._nested_eval)
    t_compile = time.perf_counter() - t0
    t0 = time.perf_counter()
    target = float(2 ** 26) if args.workload == "cfg4" else None      # cfg4 must be sliced
    plan = executor.get_plan(net, target_size=target, trials=16 if target else 8)
    t_plan = time.perf_counter() - t0
    sess = executor.Session(plan)
    sess.set_leaves([net])
    torch.cuda.synchronize()

    def step():
        sess.build_leaves()
        sess.contract(net.scalar)

    use_graph = len(plan.steps) > 8 and not plan.sliced
    if use_graph:
        sess.capture(net.scalar)
        step_fn = sess.replay
    else:
        step_fn = step
    sess.launches = 0
    step()
    launches_per_step = sess.launches

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_fn()
    barrier()
    with ClockSampler(local_rank) as clocks:
        for _ in range(args.warmup):         # the sampler start-up left the GPU idle
            step_fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step_fn()
        e1.record()
        barrier()
        secs = e0.elapsed_time(e1) * 1e-3
        # ---- dominant kernel alone, same stream, CUDA events over K launches
        hbm_bw = 6556.2e9
        try:
            hbm_bw = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get(
                "hbm_gbs", 6556.2)) * 1e9
        except OSError:
            pass
        dom = max(plan.steps, key=lambda s: max(s.bytes / hbm_bw, 8.0 * s.macs / (FP64_PIPE_TFLOPS * 1e12)))
        zeros = [0] * len(plan.sliced)
        stream = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            sess._launch(dom, zeros, 1.0 + 0j, 0, stream)
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(args.steps):
            sess._launch(dom, zeros, 1.0 + 0j, 0, stream)
        k1.record()
        torch.cuda.synchronize()
        dom_secs = k0.elapsed_time(k1) * 1e-3 / args.steps
        clocks.hold_load(step_fn, torch.cuda.synchronize)
    if world > 1:
        tt = torch.tensor([secs], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        secs = float(tt.item())
    value = world * args.steps / secs

    # ---- e2e: the user call, host objects in, host numpy out.  Measured twice: with
    # the product default (a large result is scanned on the device and, when sparse,
    # handed over as its non-zero list) and with the dense device->host copy forced.
    def time_e2e(backend):
        res = None
        for _ in range(3):                   # warm-up: plan cache + the pinned result
            res = diagram.eval(backend)      # buffers of torch's caching host allocator
        n_e2e = max(2, min(args.steps, 5))
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            d_i = factory()
            res = d_i.eval(backend)
        torch.cuda.synchronize()
        e2e = (time.perf_counter() - t0) / n_e2e
        if world > 1:
            tt = torch.tensor([e2e], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e = float(tt.item())
        return e2e, dict(B200Backend.last_d2h), res
    e2e_secs, d2h_info, res = time_e2e(B200Backend(target_size=target, trials=16 if target else 8))
    del res
    e2e_dense_secs, d2h_dense_info, res = time_e2e(
        B200Backend(target_size=target, trials=16 if target else 8, sparse_d2h=False))
    h2d = int(sess.desc_dev.numel() + sess.params_host.numel() * 16)
    d2h = int(d2h_info["bytes"])
    d2h_dense = int(res.tensor.array.nbytes)
    del res

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tensor_bound = 8.0 * dom.macs / (FP64_PIPE_TFLOPS * 1e12) > dom.bytes / (hbm_peak * 1e9)
    if tensor_bound:
        achieved, peak_val, unit = 8.0 * dom.macs / dom_secs / 1e12, FP64_PIPE_TFLOPS, "TFLOP/s"
    else:
        achieved, peak_val, unit = dom.bytes / dom_secs / 1e9, hbm_peak, "GB/s"
    roofline = {
        "bound": "tensor" if tensor_bound else "hbm", "achieved": achieved, "peak": peak_val,
        "unit": unit, "frac": achieved / peak_val,
        # dram__bytes_read+write of this kernel from `ncu --set full` on the
        # quarter-size instance (1.0179 GB written + 1.4 MB read per launch of the
        # 4-rows-per-thread kernel, profiles/r01_fma_final_step.md) scaled by 4; the
        # full-size capture does not finish under ncu's replay
        "traffic": 4.077e9 if args.workload == "cfg2" else None,
        "peak_source": ("FP64 DMMA peak measured on this pool (profiles/r01_probe.json)" if tensor_bound
                        else "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650"),
        "kernel": f"b2_contract_{dom.kernel} ({'final step, ' if dom.final else ''}(batch,M,N,K)={dom.mnk})",
        "algorithmic_bytes_per_launch": dom.bytes, "launch_ms": dom_secs * 1e3,
        "macs_per_launch": dom.macs,
    }

    # whole-plan roofline: sum over steps of max(bytes / HBM, flops / FP64 pipe)
    from optyx_b200.planner import plan_bound_seconds
    fp64_tf = FP64_PIPE_TFLOPS
    bound_s = plan_bound_seconds(plan, hbm_peak, fp64_tf)
    plan_roofline = {"bound_ms": bound_s * 1e3, "measured_ms": secs / args.steps * 1e3,
                     "frac": bound_s / (secs / args.steps), "hbm_gbs": hbm_peak,
                     "fp64_pipe_tflops": fp64_tf,
                     "note": "sum over plan steps of max(algorithmic bytes / HBM peak, 8 MACs / "
                             "FP64 DMMA peak measured on this pool, profiles/r01_probe.json)"}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": {"workload": desc, "leaves": len(net.leaves), "plan_steps": len(plan.steps),
                   "cuda_graph": bool(use_graph), "l2": "output 4.3 GB per step > 126 MB L2",
                   "plan_macs": plan.total_macs, "plan_bytes": plan.total_bytes,
                   "host_compile_ms": t_compile * 1e3, "host_plan_ms": t_plan * 1e3},
        "clocks": clocks.summary(),
        "e2e": {"value": world / e2e_secs, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_eval": e2e_secs * 1e3,
                "d2h": d2h_info.get("mode"), "result_bytes": d2h_dense,
                "note": "host diagram -> host numpy array through diagram.eval(B200Backend()); "
                        "'sparse' = the device scans the result and copies only its non-zero "
                        "(index, value) list, scattered into a zero-filled host array"},
        "e2e_dense_d2h": {"value": world / e2e_dense_secs, "unit": UNIT,
                          "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_dense,
                          "ms_per_eval": e2e_dense_secs * 1e3,
                          "note": "same call with B200Backend(sparse_d2h=False): the whole "
                                  "result crosses PCIe"},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline,
        "plan_roofline": plan_roofline,
        "contraction_tflops": 8.0 * plan.total_macs * world * args.steps / secs / 1e12,
    }

    if not args.no_sliced:
        out["sliced"] = run_sliced(world, rank)
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"], _ = cpu_reference(args.workload, 2, 0)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
    sys.stdout.flush()
    if rank == 0:
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    os.close(real_stdout)


def run_sliced(world: int, rank: int):
    """cfg4: one sliced amplitude network, slices partitioned over the ranks,
    one NCCL sum (strong scaling; time is max over ranks on the device)."""
    import torch
    import torch.distributed as dist
    from optyx_b200 import dist as b2dist
    from optyx_b200 import executor
    from optyx_b200.core.backends import B200Backend
    factory, desc = workload("cfg4")
    net = B200Backend()._get_network(factory())
    plan = executor.get_plan(net, target_size=float(2 ** 26), trials=16)
    sess = executor.Session(plan)
    sess.set_leaves([net])
    b2dist.run_sliced(sess, net)                     # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    amp = b2dist.run_sliced(sess, net)
    e1.record()
    torch.cuda.synchronize()
    secs = e0.elapsed_time(e1) * 1e-3
    if world > 1:
        tt = torch.tensor([secs], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        secs = float(tt.item())
    a = complex(amp.cpu().numpy())
    from optyx_b200.planner import plan_bound_seconds
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
    except OSError:
        hbm = 6650.0
    bound = plan_bound_seconds(plan, hbm, FP64_PIPE_TFLOPS) / world
    return {"workload": desc, "nslices": plan.nslices, "peak_elems": plan.info.peak,
            "macs": plan.total_macs, "bytes": plan.total_bytes, "ms": secs * 1e3,
            "tflops": 8.0 * plan.total_macs / secs / 1e12,
            "gbs": plan.total_bytes / secs / 1e9,
            "roofline_bound_ms": bound * 1e3, "roofline_frac": bound / secs,
            "amplitude": [a.real, a.imag], "n_gpus": world}


if __name__ == "__main__":
    main()
