#!/bin/bash
# On the GPU box: parity tests + brief bench, everything logged to gpurun_out/quick_$TAG.log.  TAG=x scripts/gpu_quick.sh [extra commands...]
TAG=${TAG:-q}
mkdir -p gpurun_out
L=gpurun_out/quick_$TAG.log
{
  if [ -z "$SKIP_TESTS" ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -4; fi
  scripts/bench_brief.sh
  for cmd in "$@"; do echo "--- $cmd"; bash -c "$cmd" 2>&1; done
} > $L 2>&1
cat $L
