import sys
sys.path.insert(0, '/root/repo/tools'); sys.path.insert(0, '/root/repo')
from bench_gemm import run
dtype = sys.argv[1] if len(sys.argv) > 1 else "c64"
layout = sys.argv[2] if len(sys.argv) > 2 else "mfast"
print(run(4096, 4096, 2048, dtype, layout, iters=1))
