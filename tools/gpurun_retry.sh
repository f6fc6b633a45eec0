#!/bin/bash
# gpurun with retries while the pod answers "transient" (no box / draining): nothing is charged for those.
# usage: tools/gpurun_retry.sh [--gpus N] --timeout S -- '<command>'
for attempt in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient"; then
    echo "[retry] attempt $attempt was transient, sleeping 150 s"; sleep 150
  else
    exit 0
  fi
done
