set -x
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest.log
(time timeout 600 python bench.py) > gpurun_out/r02_bench_n1.log 2>&1
echo "bench rc=$?" >> gpurun_out/r02_bench_n1.log
(time timeout 600 python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/r02_bench_ref.log 2>&1
for w in cfg1 cfg2 cfg3 cfg5a cfg5b; do
 (timeout 400 python bench.py --workload $w --no-cpu) > gpurun_out/r02_bench_$w.log 2>&1
done
tail -3 gpurun_out/r02_pytest.log; tail -2 gpurun_out/r02_bench_n1.log
