#!/bin/bash
# Run on the GPU box: tests, bench, launch list, ncu captures of the main kernels.  TAG=r1_h scripts/gpu_capture.sh
# Schedule of one config-2 solve (3 levels x 3 outer steps): levels 2 and 1 run irls_fused (6 launches), level 0 runs
# 18 x (pass1, pass2); 9 linearise / pose_update, 8 warp launches.  profile_step.py does two solves; the skips below
# select a finest-level instance of the second (warm) solve.
TAG=${TAG:-r1_x}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/${TAG}_pytest.log
fi
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
tail -c 3000 gpurun_out/${TAG}_bench.json
# launch list of one measured solve (second solve of profile_step; 128 pairs)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/profile_step.py 128 2 > gpurun_out/${TAG}_ls.log 2>&1
tail -1 gpurun_out/${TAG}_ls.log
for k in ${KERNELS:-irls_pass1_kernel irls_pass2_kernel linearise_kernel kmeans_kernel warp_kernel irls_fused_kernel label_connect_kernel}; do
  case $k in
    irls_pass1_kernel|irls_pass2_kernel) SKIP=${SKIP_PASS:-18};;
    linearise_kernel|pose_update_kernel) SKIP=${SKIP_LIN:-15};;
    warp_kernel|warp_normalise_kernel) SKIP=${SKIP_WARP:-13};;
    irls_fused_kernel) SKIP=${SKIP_FUSED:-9};;
    *) SKIP=1;;
  esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s $SKIP -c 1 -f -o gpurun_out/${TAG}_$k python scripts/profile_step.py 128 2 > gpurun_out/${TAG}_ncu_$k.log 2>&1
  echo "$k ncu exit $?"
done
