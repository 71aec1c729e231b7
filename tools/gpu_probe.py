"""GPU probe: FP64 pipe peaks and first kernel throughputs -> gpurun_out/probe.json"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optyx_b200 import runtime as rt  # noqa: E402

lib = rt.load()
out = {"gpu": torch.cuda.get_device_name(0), "sms": lib.b2_device_sms()}
stream = torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


sink = torch.zeros(1, dtype=torch.float64, device="cuda")
flops = C.c_double(0)
for kind, name in [(0, "dfma"), (1, "dmma_m8n8k4"), (2, "dmma_m16n8k16")]:
    t = timeit(lambda: rt.check(lib.b2_fp64_peak_kernel(kind, 4096, sink.data_ptr(), C.byref(flops), stream)))
    out[f"fp64_{name}_tflops"] = flops.value / t / 1e12


def gemm(M, N, K, kfast=True):
    A = torch.randn(M, K, dtype=torch.complex128, device="cuda")
    B = torch.randn(K, N, dtype=torch.complex128, device="cuda") if not kfast else \
        torch.randn(N, K, dtype=torch.complex128, device="cuda")
    Cc = torch.empty(M, N, dtype=torch.complex128, device="cuda")
    d = rt.GemmModes()
    d.n_batch, d.n_m, d.n_n, d.n_k = 0, 1, 1, 1
    d.dim[0], d.dim[1], d.dim[2] = M, N, K
    d.sa[0], d.sb[0], d.sc[0] = K, 0, N
    d.sa[1], d.sb[1], d.sc[1] = 0, (K if kfast else 1), 1
    d.sa[2], d.sb[2], d.sc[2] = 1, (1 if kfast else N), 0
    fn = lambda: rt.check(lib.b2_contract_gemm(C.byref(d), A.data_ptr(), B.data_ptr(), Cc.data_ptr(), 1.0, 0.0, 0, 0, stream))
    t = timeit(fn)
    ref = A @ (B.T if kfast else B)
    err = ((Cc - ref).abs().max() / ref.abs().max()).item()
    t_torch = timeit(lambda: torch.matmul(A, B.T if kfast else B))
    return {"M": M, "N": N, "K": K, "kfast": kfast, "ms": t * 1e3, "tflops": 8.0 * M * N * K / t / 1e12,
            "relerr": err, "torch_zgemm_tflops": 8.0 * M * N * K / t_torch / 1e12,
            "gbs": 16.0 * (M * K + K * N + M * N) / t / 1e9}


out["gemm"] = [gemm(4096, 4096, 4096), gemm(4096, 4096, 4096, False), gemm(8192, 8192, 512),
               gemm(1 << 16, 64, 64), gemm(1 << 20, 16, 16), gemm(2048, 2048, 2048)]


def direct_gate(nq, q):
    """apply a 2x2 gate on qubit q of an nq-qubit state (bandwidth-bound)."""
    n = 1 << nq
    A = torch.randn(n, dtype=torch.complex128, device="cuda")
    G = torch.randn(2, 2, dtype=torch.complex128, device="cuda")
    Cc = torch.empty(n, dtype=torch.complex128, device="cuda")
    hi, lo = 1 << (nq - 1 - q), 1 << q
    d = rt.Modes()
    modes = [(hi, 2 * lo, 0, 2 * lo), (2, 0, 2, lo), (lo, 1, 0, 1)]
    modes = [m for m in modes if m[0] > 1]
    d.n_out, d.n_red = len(modes), 1
    for k, m in enumerate(modes + [(2, lo, 1, 0)]):
        d.dim[k], d.sa[k], d.sb[k], d.sc[k] = m
    fn = lambda: rt.check(lib.b2_contract_direct(C.byref(d), A.data_ptr(), G.data_ptr(), Cc.data_ptr(), 1.0, 0.0, 0, 0, stream))
    t = timeit(fn)
    ref = torch.einsum("hql,nq->hnl", A.view(hi, 2, lo), G).reshape(-1)
    err = ((Cc - ref).abs().max() / ref.abs().max()).item()
    return {"nq": nq, "q": q, "ms": t * 1e3, "gbs": 32.0 * n / t / 1e9, "relerr": err}


out["direct_gate"] = [direct_gate(27, 0), direct_gate(27, 13), direct_gate(27, 26)]
x = torch.empty(1 << 27, dtype=torch.complex128, device="cuda")
y = torch.empty(1 << 27, dtype=torch.complex128, device="cuda")
t = timeit(lambda: y.copy_(x))
out["torch_copy_gbs"] = 32.0 * (1 << 27) / t / 1e9
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
print(json.dumps(out, indent=1))
