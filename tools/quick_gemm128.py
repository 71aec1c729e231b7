import sys, json
sys.path.insert(0, '/root/repo/tools'); sys.path.insert(0, '/root/repo')
from bench_gemm import run
for (M,N,K) in [(4096,4096,4096),(16384,2048,256),(32768,16384,32),(8192,32,1024)]:
    for layout in ("kfast","mfast"):
        print(json.dumps(run(M,N,K,"c128",layout)), flush=True)
