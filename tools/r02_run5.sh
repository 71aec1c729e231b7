mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q > gpurun_out/r02_test_kernels.log 2>&1; tail -5 gpurun_out/r02_test_kernels.log
python tools/plan_profile.py cfg4 > gpurun_out/r02_plan_cfg4_c.log 2>&1; head -30 gpurun_out/r02_plan_cfg4_c.log
