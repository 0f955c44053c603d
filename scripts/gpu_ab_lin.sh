#!/bin/bash
# A/B of kernel variants on one B200.  Variant libraries are built beforehand into staticfusion_b200/lib/<name>.so (git-ignored, they
# travel with the snapshot) and selected with SF_B200_LIB; "tree" is the working tree's library.  For every variant: the two main
# parity files, then bench_brief.  Everything under timeout: a hung kernel must not hold the box.
#     VARIANTS="tree ab_prev" scripts/gpu_ab_lin.sh
mkdir -p gpurun_out
for v in ${VARIANTS:-tree}; do
  echo "== $v"
  ( [ "$v" != tree ] && export SF_B200_LIB=$PWD/staticfusion_b200/lib/$v.so
    timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference.py -m gpu -x -q 2>&1 | tail -1
    STEPS=10 timeout 200 bash scripts/bench_brief.sh > gpurun_out/ab_$v.txt 2>&1
    head -1 gpurun_out/ab_$v.txt; grep "linearise_kernel L0\|linearise_kernel all\|warp stage all\|irls_iteration_finest\|irls_fused" gpurun_out/ab_$v.txt )
done
