mkdir -p gpurun_out
for thr in "32,8,8,2.0" "32,8,16,1.0"; do
B2_GEMM_THRESHOLD=$thr timeout 600 python tools/plan_profile.py cfg4 > gpurun_out/r02_plan_cfg4_thr.log 2>&1; echo "thr=$thr rc=$?"; sed -n 4,4p gpurun_out/r02_plan_cfg4_thr.log; grep -c "fma" gpurun_out/r02_plan_cfg4_thr.log
sed -n 5,30p gpurun_out/r02_plan_cfg4_thr.log | grep -v "gemm_ws  " | head -12
done
