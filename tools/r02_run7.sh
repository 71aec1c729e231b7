mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "thin_and_3m or gemm_kernel_permuted" > gpurun_out/r02_test_ws.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r02_test_ws.log
timeout 300 python tools/plan_profile.py cfg4 > gpurun_out/r02_plan_cfg4_d.log 2>&1; echo "rc=$?"; head -30 gpurun_out/r02_plan_cfg4_d.log
