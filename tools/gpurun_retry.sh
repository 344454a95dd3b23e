#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'   -- retries while the pod answers "busy / transient" (nothing charged)
T=$1; shift
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1); RC=$?
  if echo "$OUT" | grep -q "status=transient\|retry in a few minutes" || [ $RC -eq 3 ]; then
    echo "[retry $i] busy, sleeping 90 s"; sleep 90; continue
  fi
  echo "$OUT" | tail -60; exit $RC
done
echo "gave up"; exit 3
