mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "thin_and_3m or gemm_kernel_permuted or random_networks" > gpurun_out/r02_test_kernels_f.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02_test_kernels_f.log
timeout 600 python tools/plan_profile.py cfg4 > gpurun_out/r02_plan_cfg4_h.log 2>&1; echo "rc=$?"; head -18 gpurun_out/r02_plan_cfg4_h.log
