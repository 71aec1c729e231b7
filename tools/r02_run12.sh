mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_baseline_sizes.py -x -q -k "not cfg4") > gpurun_out/r02_pytest_e.log 2>&1; tail -3 gpurun_out/r02_pytest_e.log
for w in cfg5a cfg5b cfg3; do
 (timeout 400 python bench.py --workload $w --no-cpu) > gpurun_out/r02_bench_${w}_e.log 2>&1
 grep -E '^\{' gpurun_out/r02_bench_${w}_e.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'])"
done
python tools/plan_profile.py cfg5a > gpurun_out/r02_plan_cfg5a_b.log 2>&1; head -12 gpurun_out/r02_plan_cfg5a_b.log
