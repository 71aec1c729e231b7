"""Debug driver for the complex64 tcgen05 GEMM: one tile, structured inputs."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optyx_b200 import runtime as rt  # noqa: E402

lib = rt.load()
M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (128, 64, 16)
rng = np.random.default_rng(0)
A = (rng.standard_normal((M, K)) + 1j * rng.standard_normal((M, K))).astype(np.complex64)
B = (rng.standard_normal((K, N)) + 1j * rng.standard_normal((K, N))).astype(np.complex64)
desc = rt.GemmModes()
desc.n_batch, desc.n_m, desc.n_n, desc.n_k = 0, 1, 1, 1
desc.dim[0], desc.sa[0], desc.sb[0], desc.sc[0] = M, K, 0, N
desc.dim[1], desc.sa[1], desc.sb[1], desc.sc[1] = N, 0, 1, 1
desc.dim[2], desc.sa[2], desc.sb[2], desc.sc[2] = K, 1, N, 0
dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
dC = torch.zeros((M, N), dtype=torch.complex64, device="cuda")
rt.check(lib.b2_contract_gemm(C.byref(desc), dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), 1.0, 0.0,
                              0, rt.B2_C64, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
got = dC.cpu().numpy().astype(np.complex128)
want = A.astype(np.complex128) @ B.astype(np.complex128)
err = np.abs(got - want)
print("max err", err.max(), "scale", np.abs(want).max())
if err.max() > 1e-4:
    print("got[0,:4]", got[0, :4])
    print("want[0,:4]", want[0, :4])
    bad = np.argwhere(err > 1e-4)
    print("bad count", len(bad), "first", bad[:5].tolist(), "rows bad", np.unique(bad[:, 0])[:20],
          "cols bad", np.unique(bad[:, 1])[:20])
