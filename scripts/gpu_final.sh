#!/bin/bash
# Closing run of a change on one B200: full GPU suite, smoke, the default bench line, one ncu capture of KERNEL and the launch
# list of the bench command.  TAG=r2w KERNEL=linearise_kernel SKIP=15 scripts/gpu_final.sh
TAG=${TAG:-r2w}; KERNEL=${KERNEL:-linearise_kernel}; SKIP=${SKIP:-15}
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json | cut -c1-600
SF_LANES=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$KERNEL" -s $SKIP -c 1 -f -o gpurun_out/${TAG}_$KERNEL python scripts/profile_step.py 512 2 > gpurun_out/${TAG}_ncu_$KERNEL.log 2>&1; echo "ncu rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
