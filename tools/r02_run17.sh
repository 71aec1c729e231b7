mkdir -p gpurun_out
(timeout 900 python bench.py --no-also) > gpurun_out/r02_bench_n1_f.log 2> gpurun_out/r02_bench_n1_f.err; echo "rc=$?"
grep -E '^\{' gpurun_out/r02_bench_n1_f.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_eval'], 'plan', d['plan_roofline'], 'TF', d['contraction_tflops'])
print(d['roofline']); print(d['config']); print(d['cpu_baseline']); print(d['result_head'], d['gpu_launches'])
"
tail -3 gpurun_out/r02_bench_n1_f.err
