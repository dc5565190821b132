#!/bin/bash
# gpurun wrapper: retries while the pod answers "transient" (busy, nothing charged).  usage: tools/gpurun_retry.sh <timeout_s> '<command>'
T=$1; shift
for attempt in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  echo "$out" | tail -40
  exit 0
done
echo "gpurun: still busy after 30 attempts"
