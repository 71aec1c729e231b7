# usage: bash tools/r02_run_n.sh N
N=$1
mkdir -p gpurun_out
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3) > gpurun_out/r02_bench_n$N.log 2> gpurun_out/r02_bench_n$N.err
echo "rc=$?"
grep -E '^\{' gpurun_out/r02_bench_n$N.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'scaling', d['scaling'], 'plan', d['plan_roofline']['frac'], 'TF', d['contraction_tflops'])
print(d.get('replicas')); print(d['result_head'], d['clocks'])
"
grep -i "nccl info.*nranks\|NCCL INFO comm.*nranks" gpurun_out/r02_bench_n$N.err | head -3
tail -3 gpurun_out/r02_bench_n$N.err
