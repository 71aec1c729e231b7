mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r02_pytest_b.log 2>&1; tail -3 gpurun_out/r02_pytest_b.log
(timeout 600 python bench.py --no-cpu) > gpurun_out/r02_bench_n1_b.log 2>&1; tail -c 600 gpurun_out/r02_bench_n1_b.log
