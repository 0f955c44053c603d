#!/bin/bash
# Run on the GPU box: tests, both bench arms, launch lists, ncu captures of the main kernels.  TAG=r1_q scripts/gpu_capture.sh
# The ncu --set full captures run the schedule on ONE stream (SF_LANES=1), so launch indices are fixed: one config-2 solve
# (3 levels x 3 outer steps) has 9 linearise / pose_update, 8 warp, 6 irls_fused (levels 2 and 1) and 3 irls_loop (level 0)
# launches; profile_step.py does two solves and the skips below select a finest-level instance of the second (warm) one.
TAG=${TAG:-r1_x}
PAIRS=${PAIRS:-512}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/${TAG}_pytest.log
fi
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "reference arm exit $?"
# launch list of the bench command itself (three-lane graph replays; ncu serialises the kernels) and of one measured solve
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1; echo "bench launch list exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/profile_step.py $PAIRS 2 > gpurun_out/${TAG}_ls.log 2>&1
tail -1 gpurun_out/${TAG}_ls.log
export SF_LANES=1
for k in ${KERNELS:-irls_loop_kernel linearise_kernel kmeans_kernel warp_kernel warp_normalise_kernel irls_fused_kernel label_connect_kernel}; do
  case $k in
    irls_loop_kernel) SKIP=${SKIP_LOOP:-3};;
    irls_pass1_kernel|irls_pass2_kernel) SKIP=${SKIP_PASS:-18};;
    linearise_kernel|pose_update_kernel) SKIP=${SKIP_LIN:-15};;
    warp_kernel|warp_normalise_kernel) SKIP=${SKIP_WARP:-13};;
    irls_fused_kernel) SKIP=${SKIP_FUSED:-9};;
    *) SKIP=1;;
  esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$k|^void $k" -s $SKIP -c 1 -f -o gpurun_out/${TAG}_$k python scripts/profile_step.py $PAIRS 2 > gpurun_out/${TAG}_ncu_$k.log 2>&1
  echo "$k ncu exit $?"
done
