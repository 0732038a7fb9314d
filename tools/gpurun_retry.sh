#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun flags] -- '<command>'   (retries while the pod answers "transient"/busy)
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -15
  if echo "$out" | grep -q "status=transient\|rc=3\b"; then sleep 90; continue; fi
  break
done
