mkdir -p gpurun_out
for s in 0 32 64; do
B2_GEMM_3M=$s timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm or random_networks" > gpurun_out/r02_test_gemm_$s.log 2>&1; tail -2 gpurun_out/r02_test_gemm_$s.log
done
python tools/bench_gemm3m.py > gpurun_out/r02_gemm3m_v2.log 2>&1; tail -5 gpurun_out/r02_gemm3m_v2.log
