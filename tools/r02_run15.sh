mkdir -p gpurun_out
for c in 16 8 4 2; do echo "ctas_per_sm=$c"; B2_LEAF_CTAS_PER_SM=$c python tools/prof_dominant.py cfg5a leaf 2>&1 | tail -1; done
