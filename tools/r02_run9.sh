mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q > gpurun_out/r02_test_kernels_d.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02_test_kernels_d.log
timeout 300 python tools/plan_profile.py cfg4 > gpurun_out/r02_plan_cfg4_e.log 2>&1; echo "rc=$?"; head -34 gpurun_out/r02_plan_cfg4_e.log
timeout 300 python tools/bench_gemm.py > gpurun_out/r02_bench_gemm.log 2>&1; grep c64 gpurun_out/r02_bench_gemm.log
