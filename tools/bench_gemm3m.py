"""3M vs 4-product FP64 DMMA contraction kernel: accuracy against torch.matmul (c128) and
algorithmic TFLOP/s (8*M*N*K / t), one subprocess per B2_GEMM_3M setting (0 | 32 | 64)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPES = [(4096, 4096, 4096), (16384, 512, 2048), (4096, 1024, 4096), (4096, 512, 2048),
          (1024, 256, 8192), (16384, 2048, 1024), (8192, 8192, 512), (16384, 2048, 256), (2048, 2048, 128)]

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import ctypes as C
    import torch
    from optyx_b200 import runtime as rt
    from tools.bench_gemm import run
    lib = rt.load()
    stream = torch.cuda.current_stream().cuda_stream
    # accuracy: kfast layout, C = A @ B^T
    M, N, K = 512, 384, 1024
    A = torch.randn(M, K, dtype=torch.complex128, device="cuda")
    B = torch.randn(N, K, dtype=torch.complex128, device="cuda")
    Cc = torch.empty(M, N, dtype=torch.complex128, device="cuda")
    d = rt.GemmModes()
    d.n_batch, d.n_m, d.n_n, d.n_k = 0, 1, 1, 1
    d.dim[0], d.sa[0], d.sb[0], d.sc[0] = M, K, 0, N
    d.dim[1], d.sa[1], d.sb[1], d.sc[1] = N, 0, K, 1
    d.dim[2], d.sa[2], d.sb[2], d.sc[2] = K, 1, 1, 0
    rt.check(lib.b2_contract_gemm(C.byref(d), A.data_ptr(), B.data_ptr(), Cc.data_ptr(), 1.0, 0.0, 0,
                                  rt.B2_C128, stream))
    want = A @ B.T
    err = float((Cc - want).abs().max() / want.abs().max())
    print(json.dumps({"setting": os.environ.get("B2_GEMM_3M"), "kernel": lib.b2_last_kernel().decode(),
                      "max_rel_err": err}), flush=True)
    for (M, N, K) in SHAPES:
        for layout in ("kfast", "mfast"):
            r = run(M, N, K, "c128", layout)
            r["setting"] = os.environ.get("B2_GEMM_3M")
            print(json.dumps(r), flush=True)
else:
    for setting in ("0", "64", "32"):
        env = dict(os.environ, B2_GEMM_3M=setting)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, cwd=ROOT)
