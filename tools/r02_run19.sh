mkdir -p gpurun_out
B2_PROFILE_TIMED=cfg4 python bench.py --steps 1 --warmup 3 --no-also --no-cpu > gpurun_out/plain_cfg4.log 2>&1 &&
B2_PROFILE_TIMED=cfg4 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r02_launches_cfg4_b.csv python bench.py --steps 1 --warmup 3 --no-also --no-cpu > gpurun_out/ncu_cfg4.log 2>&1
echo "rc=$?"; tail -c 400 gpurun_out/plain_cfg4.log
