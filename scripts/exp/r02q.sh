#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -q -m gpu -x -k "chain" > gpurun_out/r02q_chain.log 2>&1
echo "chain exit=$?"; tail -n 6 gpurun_out/r02q_chain.log
timeout 300 python scripts/trace_chain.py 3 6 1100 > gpurun_out/r02q_trace_b3.log 2>&1; echo "trace exit=$?"; head -n 12 gpurun_out/r02q_trace_b3.log
timeout 300 python scripts/trace_chain.py 24 6 1100 > gpurun_out/r02q_trace_b24.log 2>&1; echo "trace exit=$?"; head -n 7 gpurun_out/r02q_trace_b24.log
OCRB_CHAIN_MAX_B=0 timeout 300 python scripts/trace_chain.py 24 6 1100 2>&1 | grep "per layer"
