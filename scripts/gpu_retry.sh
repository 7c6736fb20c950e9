#!/bin/bash
# usage: scripts/gpu_retry.sh <log> <timeout_s> <command...>   -- retries a gpurun call while the pod answers busy (exit 3)
log=$1; shift; to=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|nothing was charged" "$log" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "gpu_retry done rc=$rc" >> "$log"
