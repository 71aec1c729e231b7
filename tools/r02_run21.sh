mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_eval.py tests/test_gpu_kernels.py -x -q) > gpurun_out/r02_pytest_g.log 2>&1; tail -2 gpurun_out/r02_pytest_g.log
(timeout 900 python bench.py --no-cpu) > gpurun_out/r02_bench_n1_g.log 2> gpurun_out/r02_bench_n1_g.err; echo rc=$?
grep -E '^\{' gpurun_out/r02_bench_n1_g.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_eval'])
print('cfg2', d['cfg2']['value'], d['cfg2']['e2e']['value'], d['cfg2']['e2e']['ms_per_eval'])
"
