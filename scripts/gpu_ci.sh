#!/bin/bash
# Runs on the GPU box (under gpurun): each GPU test file in its own process with a timeout, so a
# trapped kernel cannot take the other files down.  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in "$@"; do
  name=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -q -s -m gpu > "gpurun_out/$name.log" 2>&1
  code=$?
  echo "$name exit=$code" | tee -a gpurun_out/ci_summary.txt
  tail -n 25 "gpurun_out/$name.log"
  [ $code -ne 0 ] && rc=1
done
exit $rc
