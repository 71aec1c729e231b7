"""Short-K dense pairs: the ctile kernel vs the classic 64x64 DMMA tiling
(B2_CTILE_KMAX decides which one b2_contract_gemm takes)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_gemm
    for (M, N, K) in [(16384, 1024, 32), (32768, 1024, 32), (32768, 16384, 64), (16384, 2048, 16),
                      (65536, 256, 64), (4096, 4096, 4096), (1024, 1024, 32768)]:
        for layout in ("kfast", "mfast"):
            print(json.dumps(bench_gemm.run(M, N, K, "c128", layout)), flush=True)
else:
    for kmax in ("32", "0"):
        env = dict(os.environ, B2_CTILE_KMAX=kmax)
        print("B2_CTILE_KMAX =", kmax, flush=True)
        subprocess.run([sys.executable, __file__, "child"], env=env)
