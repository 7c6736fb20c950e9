#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -q -m gpu -x -k "skinny or chain or plan" > gpurun_out/r02ag_dense.log 2>&1; echo "dense tests exit=$?"; tail -n 4 gpurun_out/r02ag_dense.log
timeout 600 python -m pytest tests/test_gpu_vlm.py -q -m gpu -x > gpurun_out/r02ag_vlm.log 2>&1; echo "vlm tests exit=$?"; tail -n 3 gpurun_out/r02ag_vlm.log
for B in 3 24 63 96; do echo "B=$B"; timeout 300 python scripts/trace_chain.py $B 4 1100 2>&1 | grep "per layer"; done
