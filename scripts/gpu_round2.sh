#!/bin/bash
# On a GPU box with N GPUs: GPU tests, bench at N = 1 and N = all, config 4 sharded.  TAG=r2b NG=2 scripts/gpu_round2.sh
TAG=${TAG:-r2}
NG=${NG:-$(nvidia-smi -L | wc -l)}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then python -m pytest ${TESTS:-tests} -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest.log; fi
if [ -z "$SKIP_N1" ]; then python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench n1 exit $?"; fi
tail -3 gpurun_out/${TAG}_bench_n1.err
if [ "$NG" -gt 1 ]; then
  for cfg in ${CFGS:-2 4}; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $NG --steps ${STEPS:-10} --warmup 3 --config $cfg --no-cpu-baseline > gpurun_out/${TAG}_bench_c${cfg}_n${NG}.json 2> gpurun_out/${TAG}_bench_c${cfg}_n${NG}.err; echo "bench config $cfg n$NG exit $?"
    tail -3 gpurun_out/${TAG}_bench_c${cfg}_n${NG}.err
  done
fi
for f in gpurun_out/${TAG}_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print(sys.argv[1], 'unreadable', e); sys.exit()
r=d['roofline']
print(sys.argv[1].split('/')[-1], 'n', d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'fps', round(d['frames_per_s']), 'iters/s', round(d['value']), 'roof', r['frac'], 'whole', r['whole_step']['frac'], 'e2e fps', round(d['e2e']['frames_per_s']), 'e2e ok', d['e2e']['matches_device_run_bitwise'], 'gather ok', d['run']['gathered_table_matches_local_rows'], d['clocks']['reasons'])
PY
done
