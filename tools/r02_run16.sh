mkdir -p gpurun_out
timeout 600 python tools/plan_profile.py cfg4 > gpurun_out/r02_plan_cfg4_g.log 2>&1; echo "rc=$?"; head -44 gpurun_out/r02_plan_cfg4_g.log
(timeout 900 python -m pytest tests/test_gpu_baseline_sizes.py -x -q -k "cfg4") > gpurun_out/r02_pytest_cfg4.log 2>&1; tail -3 gpurun_out/r02_pytest_cfg4.log
