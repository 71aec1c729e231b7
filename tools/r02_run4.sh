mkdir -p gpurun_out
export B2_GEMM_3M=32
for l in mfast kfast; do
python tools/prof_gemm.py c128 $l > gpurun_out/plain_g3m_$l.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm3m -c 1 -f -o gpurun_out/r02_gemm3m_$l python tools/prof_gemm.py c128 $l > gpurun_out/ncu_g3m_$l.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
