mkdir -p gpurun_out
(timeout 1700 python -m pytest tests -m gpu -x -q) > gpurun_out/r02_pytest_final.log 2>&1; tail -3 gpurun_out/r02_pytest_final.log
(timeout 900 python bench.py) > gpurun_out/r02_bench_n1_final.log 2> gpurun_out/r02_bench_n1_final.err; echo "bench rc=$?"
(timeout 600 python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/r02_bench_ref_final.log 2>&1
for w in cfg1 cfg2 cfg3 cfg5a cfg5b; do
 (timeout 400 python bench.py --workload $w) > gpurun_out/r02_bench_${w}_final.log 2>&1
done
python tools/prof_dominant.py cfg4 > gpurun_out/r02_dom_cfg4_final.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -c 3 -f -o gpurun_out/r02_dom_cfg4_final \
   python tools/prof_dominant.py cfg4 > gpurun_out/ncu_dom_cfg4_final.log 2>&1
grep key= gpurun_out/r02_dom_cfg4_final.log
