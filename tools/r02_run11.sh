mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02_pytest_d.log 2>&1; tail -3 gpurun_out/r02_pytest_d.log
(timeout 600 python bench.py) > gpurun_out/r02_bench_n1_d.log 2>&1; tail -c 300 gpurun_out/r02_bench_n1_d.log
for w in cfg1 cfg2 cfg3 cfg5a cfg5b; do
 (timeout 400 python bench.py --workload $w) > gpurun_out/r02_bench_${w}_d.log 2>&1
done
python tools/plan_profile.py cfg3 > gpurun_out/r02_plan_cfg3_c.log 2>&1; head -12 gpurun_out/r02_plan_cfg3_c.log
