#!/bin/bash
# On the GPU box: compute-sanitizer memcheck + racecheck (+ synccheck) over scripts/sanitize_batch.py; logs to gpurun_out/${TAG}_sanitizer_*.log
TAG=${TAG:-r2}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_batch.py > gpurun_out/${TAG}_sanitizer_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_batch ok|smoke ok" gpurun_out/${TAG}_sanitizer_$tool.log
done
