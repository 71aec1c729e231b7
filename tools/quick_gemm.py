import sys
sys.path.insert(0, '/root/repo/tools'); sys.path.insert(0, '/root/repo')
import json
from bench_gemm import run
for (M,N,K) in [(4096,4096,4096),(8192,896,16384),(16384,2048,256)]:
    for layout in ("kfast","mfast"):
        print(json.dumps(run(M,N,K,"c64",layout)), flush=True)
