#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3): scripts/gpu_retry.sh [gpurun args] -- 'command'
for attempt in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpu_retry] busy, attempt $attempt; sleeping 120 s" >&2
  sleep 120
done
exit 3
