#!/bin/bash
# On a 1-GPU box: bench configs 3, 4 (whole sequence on one GPU) and 5, the reference arm and the CPU baseline script.  TAG=r2h scripts/gpu_configs.sh
TAG=${TAG:-r2}
mkdir -p gpurun_out
for cfg in ${CONFIGS:-3 4 5}; do
  python bench.py --config $cfg --steps ${STEPS:-6} --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_config${cfg}.json 2> gpurun_out/${TAG}_bench_config${cfg}.err; echo "config $cfg exit $?"; tail -2 gpurun_out/${TAG}_bench_config${cfg}.err
done
if [ -z "$SKIP_CPU" ]; then
python scripts/cpu_baseline.py --config 2 --pairs 24 --out gpurun_out/${TAG}_cpu_baseline_config2.json > /dev/null 2> gpurun_out/${TAG}_cpu_baseline.err; echo "cpu baseline exit $?"
fi
for f in gpurun_out/${TAG}_bench_config*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print(sys.argv[1], 'unreadable', e); sys.exit()
r=d['roofline']
print(sys.argv[1].split('/')[-1], 'ms/step', round(d['ms_per_step'],3), 'fps', round(d['frames_per_s']), 'iters/s', round(d['value']), 'roof', r['frac'], 'whole', r['whole_step']['frac'], 'e2e fps', round(d['e2e']['frames_per_s']), 'e2e ok', d['e2e']['matches_device_run_bitwise'], d['clocks']['reasons'])
PY
done
