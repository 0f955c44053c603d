#!/bin/bash
# gpurun with retries while the pod answers busy (exit code 3 / "transient"): scripts/gpurun_retry.sh [gpurun args...]
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "gpurun_retry: gave up"; exit 3
