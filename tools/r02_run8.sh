mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02_pytest_c.log 2>&1; tail -3 gpurun_out/r02_pytest_c.log
(timeout 600 python bench.py) > gpurun_out/r02_bench_n1_c.log 2>&1; tail -c 300 gpurun_out/r02_bench_n1_c.log
python tools/prof_dominant.py cfg4 > gpurun_out/r02_dom_cfg4_ws.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -c 3 -f -o gpurun_out/r02_dom_cfg4_ws \
   python tools/prof_dominant.py cfg4 > gpurun_out/ncu_dom_cfg4_ws.log 2>&1
python tools/plan_profile.py cfg3 > gpurun_out/r02_plan_cfg3_b.log 2>&1
python tools/plan_profile.py cfg1 > gpurun_out/r02_plan_cfg1_b.log 2>&1
