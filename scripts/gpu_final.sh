#!/bin/bash
# Closing run of a change on one B200: full GPU suite, smoke, the default bench line; optionally (AB_LIB=file in
# staticfusion_b200/lib) a short bench of a previous build on the same box, an ncu capture of KERNEL and the launch list of the
# bench command (NCU=1).  TAG=r2x scripts/gpu_final.sh
TAG=${TAG:-r2x}; KERNEL=${KERNEL:-linearise_kernel}; SKIP=${SKIP:-15}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 200 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-240 gpurun_out/${TAG}_bench.json
if [ -n "$AB_LIB" ]; then
  SF_B200_LIB=$PWD/staticfusion_b200/lib/$AB_LIB STEPS=10 timeout 120 bash scripts/bench_brief.sh > gpurun_out/${TAG}_ab_previous.txt 2>&1; head -1 gpurun_out/${TAG}_ab_previous.txt; grep "linearise_kernel" gpurun_out/${TAG}_ab_previous.txt
fi
if [ -n "$NCU" ]; then
  SF_LANES=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$KERNEL" -s $SKIP -c 1 -f -o gpurun_out/${TAG}_$KERNEL python scripts/profile_step.py 512 2 > gpurun_out/${TAG}_ncu_$KERNEL.log 2>&1; echo "ncu rc=$?"
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
fi
