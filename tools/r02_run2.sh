set -x
mkdir -p gpurun_out
python tools/plan_profile.py cfg4 > gpurun_out/r02_plan_cfg4.log 2>&1
python tools/plan_profile.py cfg3 > gpurun_out/r02_plan_cfg3.log 2>&1
python tools/plan_profile.py cfg1 > gpurun_out/r02_plan_cfg1.log 2>&1
python tools/plan_profile.py cfg5a > gpurun_out/r02_plan_cfg5a.log 2>&1
# launch lists of the timed region of bench.py (one step)
B2_PROFILE_TIMED=cfg4 python bench.py --steps 1 --warmup 3 --no-also --no-cpu > gpurun_out/plain_cfg4.log 2>&1 &&
B2_PROFILE_TIMED=cfg4 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r02_launches_cfg4.csv python bench.py --steps 1 --warmup 3 --no-also --no-cpu > gpurun_out/ncu_cfg4.log 2>&1
B2_PROFILE_TIMED=cfg2 python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_cfg2.log 2>&1 &&
B2_PROFILE_TIMED=cfg2 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r02_launches_cfg2.csv python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_cfg2.log 2>&1
# full captures of the dominant kernels
for w in cfg4 cfg2; do
python tools/prof_dominant.py $w > gpurun_out/r02_dom_$w.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -c 3 -f -o gpurun_out/r02_dom_$w \
   python tools/prof_dominant.py $w > gpurun_out/ncu_dom_$w.log 2>&1
done
python tools/prof_dominant.py cfg5a leaf > gpurun_out/r02_leaf_cfg5a.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -c 3 -f -o gpurun_out/r02_leaf_cfg5a \
   python tools/prof_dominant.py cfg5a leaf > gpurun_out/ncu_leaf.log 2>&1
python tools/prof_dominant.py cfg4 small > gpurun_out/r02_small_cfg4.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -c 3 -f -o gpurun_out/r02_small_cfg4 \
   python tools/prof_dominant.py cfg4 small > gpurun_out/ncu_small.log 2>&1
ls -la gpurun_out
