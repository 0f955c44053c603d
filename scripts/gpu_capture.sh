#!/bin/bash
# Run on the GPU box: tests, bench, launch list, ncu captures of the main kernels.  TAG=r1_g scripts/gpu_capture.sh
TAG=${TAG:-r1_x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" 
tail -3 gpurun_out/${TAG}_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
tail -c 3000 gpurun_out/${TAG}_bench.json
# launch list of one measured solve (second solve of profile_step; 128 pairs)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/profile_step.py 128 2 > gpurun_out/${TAG}_ls.log 2>&1
tail -1 gpurun_out/${TAG}_ls.log
for k in irls_pass1_kernel irls_pass2_kernel linearise_kernel kmeans_kernel warp_kernel pose_update_kernel; do
  # skip the whole first solve and the coarser levels: take the LAST-but-few instance by skipping many
  case $k in
    irls_pass1_kernel|irls_pass2_kernel) SKIP=${SKIP_PASS:-92};;
    linearise_kernel|warp_kernel|pose_update_kernel) SKIP=${SKIP_LIN:-15};;
    kmeans_kernel) SKIP=1;;
  esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$k" -s $SKIP -c 1 -f -o gpurun_out/${TAG}_$k python scripts/profile_step.py 128 2 > gpurun_out/${TAG}_ncu_$k.log 2>&1
  echo "$k ncu exit $?"
done
