mkdir -p gpurun_out
for l in kfast; do
python tools/prof_gemm.py c64 $l > gpurun_out/plain_c64_$l.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_c64 -c 1 -f -o gpurun_out/r02_c64_$l python tools/prof_gemm.py c64 $l > gpurun_out/ncu_c64_$l.log 2>&1
done
cat gpurun_out/plain_c64_kfast.log | tail -1
timeout 300 python tools/plan_profile.py cfg5a > gpurun_out/r02_plan_cfg5a_c.log 2>&1; head -8 gpurun_out/r02_plan_cfg5a_c.log
