# the round's final validation on one B200: smoke, the whole GPU suite, the default bench line
mkdir -p gpurun_out
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
(timeout 1700 python -m pytest tests -m gpu -x -q) > gpurun_out/r02_pytest_final.log 2>&1; tail -2 gpurun_out/r02_pytest_final.log
(timeout 900 python bench.py) > gpurun_out/r02_bench_n1_final.log 2> gpurun_out/r02_bench_n1_final.err; echo "bench rc=$?"
