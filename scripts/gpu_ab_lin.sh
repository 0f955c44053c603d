#!/bin/bash
# A/B of linearise / warp kernel variants on one B200.  Variant libraries are built beforehand into staticfusion_b200/lib/ab_*.so
# and selected with SF_B200_LIB; the default library is the working tree's.  Everything under timeout.
mkdir -p gpurun_out
run() {  # name, lib ("" = default)
  echo "== $1"
  ( [ -n "$2" ] && export SF_B200_LIB=$PWD/staticfusion_b200/lib/$2
    timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference.py -m gpu -x -q 2>&1 | tail -1
    STEPS=10 timeout 200 bash scripts/bench_brief.sh > gpurun_out/ab_$1.txt 2>&1
    head -1 gpurun_out/ab_$1.txt; grep "linearise_kernel L0\|linearise_kernel all\|warp stage all\|irls_iteration_finest\|irls_fused" gpurun_out/ab_$1.txt )
}
run v2 ""
run v3 ab_v3.so
run head ab_head.so
export SF_LANES=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^linearise_kernel|^void sf::linearise_kernel|linearise_kernel" -s 15 -c 1 -f -o gpurun_out/r2w_linearise_kernel python scripts/profile_step.py 512 2 > gpurun_out/r2w_ncu_linearise.log 2>&1; echo "ncu exit $?"
