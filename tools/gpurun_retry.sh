#!/bin/bash
# Usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'   -- retries while the pod answers "busy / draining"
T=$1; shift
for attempt in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then
    echo "[retry $attempt] pod busy, sleeping 90 s" >&2
    sleep 90
    continue
  fi
  echo "$out"
  exit 0
done
echo "gave up after 30 attempts"
exit 3
