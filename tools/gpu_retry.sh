#!/bin/bash
# usage: tools/gpu_retry.sh <log> <gpurun args...>   — retries while the pod answers "transient" (busy)
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient\|rc=3" "$log"; then sleep 90; else break; fi
done
