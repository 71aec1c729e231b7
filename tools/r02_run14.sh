mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "c64 or complex64" > gpurun_out/r02_test_c64.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02_test_c64.log
timeout 300 python tools/bench_gemm.py > gpurun_out/r02_bench_gemm_b.log 2>&1; grep c64 gpurun_out/r02_bench_gemm_b.log
