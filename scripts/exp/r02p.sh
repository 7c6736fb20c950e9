#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/trace_chain.py 3 6 1100 > gpurun_out/r02p_trace_b3.log 2>&1; echo "trace exit=$?"; head -n 40 gpurun_out/r02p_trace_b3.log
OCRB_CHAIN_MAX_B=0 timeout 300 python scripts/trace_chain.py 3 6 1100 2>&1 | grep "per layer"
timeout 300 python scripts/trace_chain.py 96 6 1100 > gpurun_out/r02p_trace_b96.log 2>&1; echo "trace exit=$?"; head -n 22 gpurun_out/r02p_trace_b96.log
OCRB_CHAIN_MAX_B=0 timeout 300 python scripts/trace_chain.py 96 6 1100 2>&1 | grep "per layer"
