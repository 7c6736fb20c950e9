#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -q -m gpu -x -k "decode_attention or plan" > gpurun_out/r02ab_attn.log 2>&1; echo "attention tests exit=$?"; tail -n 4 gpurun_out/r02ab_attn.log
for B in 3 24 96; do echo "B=$B"; timeout 300 python scripts/trace_chain.py $B 4 1100 2>&1 | grep "per layer"; done
