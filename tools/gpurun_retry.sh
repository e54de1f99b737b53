#!/bin/bash
# gpurun with retries while the pod answers "busy / transient" (exit 3 or status=transient): nothing is charged for those.
# usage: tools/gpurun_retry.sh <timeout_s> [--gpus N] -- '<command>'
T=$1; shift
for try in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(/usr/local/graft/bin/gpurun --timeout $T "$@" 2>&1); rc=$?
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|retry in a few minutes\|no box or slot" || [ $rc = 3 ]; then
    echo "[retry $try] busy, sleeping 150 s"; sleep 150; continue
  fi
  exit $rc
done
exit 3
