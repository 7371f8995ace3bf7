#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> '<command>' [gpus]   -- retries while the pool answers "busy / draining"
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 30); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$CMD" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$CMD" 2>&1); fi
  RC=$?
  if echo "$OUT" | grep -q "status=transient\|nothing was charged"; then echo "[retry $i] transient, sleeping"; sleep 90; continue; fi
  if [ $RC -eq 3 ]; then echo "[retry $i] rc=3, sleeping"; sleep 90; continue; fi
  echo "$OUT" | tail -40; exit $RC
done
echo "gave up"; exit 3
