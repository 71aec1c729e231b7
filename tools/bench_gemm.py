"""Micro-benchmark of the two tensor-core contraction kernels behind
b2_contract_gemm: complex128 (FP64 DMMA) and complex64 (tcgen05 kind::tf32,
3-product split).  Reports algorithmic TFLOP/s = 8*B*M*N*K / t (CUDA events)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optyx_b200 import runtime as rt  # noqa: E402

lib = rt.load()
stream = torch.cuda.current_stream().cuda_stream


def run(M, N, K, dtype, layout="kfast", iters=5):
    """layout kfast: A[m,k], B[n,k] (K contiguous in both); mfast: A[k,m], B[k,n]."""
    tdt = torch.complex128 if dtype == "c128" else torch.complex64
    code = rt.B2_C128 if dtype == "c128" else rt.B2_C64
    A = torch.randn(M * K, dtype=tdt, device="cuda")
    B = torch.randn(N * K, dtype=tdt, device="cuda")
    Cc = torch.empty(M * N, dtype=tdt, device="cuda")
    d = rt.GemmModes()
    d.n_batch, d.n_m, d.n_n, d.n_k = 0, 1, 1, 1
    if layout == "kfast":
        sa_m, sa_k, sb_n, sb_k = K, 1, K, 1
    else:
        sa_m, sa_k, sb_n, sb_k = 1, M, 1, N
    d.dim[0], d.sa[0], d.sb[0], d.sc[0] = M, sa_m, 0, N
    d.dim[1], d.sa[1], d.sb[1], d.sc[1] = N, 0, sb_n, 1
    d.dim[2], d.sa[2], d.sb[2], d.sc[2] = K, sa_k, sb_k, 0

    def call():
        rt.check(lib.b2_contract_gemm(C.byref(d), A.data_ptr(), B.data_ptr(), Cc.data_ptr(), 1.0, 0.0,
                                      0, code, stream))
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 8.0 * M * N * K / ms / 1e9
    es = 16 if dtype == "c128" else 8
    gbs = es * (M * K + N * K + M * N) / ms / 1e6
    return {"dtype": dtype, "layout": layout, "M": M, "N": N, "K": K, "ms": round(ms, 4),
            "tflops": round(tf, 2), "gbs": round(gbs, 1), "kernel": lib.b2_last_kernel().decode()}


if __name__ == "__main__":
    out = []
    shapes = [(4096, 4096, 4096), (8192, 896, 16384), (16384, 2048, 256), (32768, 16384, 64),
              (8192, 8192, 512)]
    for dtype in ("c64", "c128"):
        for (M, N, K) in shapes:
            if dtype == "c128" and M * N * K > 2 ** 37:
                continue
            for layout in ("kfast", "mfast"):
                r = run(M, N, K, dtype, layout)
                out.append(r)
                print(json.dumps(r), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bench_gemm.json"), "w") as fh:
        json.dump(out, fh, indent=1)
