#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -q -m gpu -x -k "chain or plan" > gpurun_out/r02s_chain.log 2>&1
echo "chain exit=$?"; tail -n 5 gpurun_out/r02s_chain.log
OCRB_CHAIN_PREFETCH=0 timeout 300 python scripts/trace_chain.py 3 4 1100 > gpurun_out/r02s_trace_b3_pf0.log 2>&1; echo "trace exit=$?"; head -n 16 gpurun_out/r02s_trace_b3_pf0.log
for PF in 8 16 24 40; do echo "PF=$PF"; OCRB_CHAIN_PREFETCH=$PF timeout 300 python scripts/trace_chain.py 3 4 1100 2>&1 | grep "per layer"; done
timeout 300 python scripts/trace_chain.py 3 4 1100 > gpurun_out/r02s_trace_b3_pf24.log 2>&1; sed -n 9,15p gpurun_out/r02s_trace_b3_pf24.log
OCRB_CHAIN_MAX_B=0 timeout 300 python scripts/trace_chain.py 3 4 1100 2>&1 | grep "per layer"
for B in 24 96; do echo "B=$B"; OCRB_CHAIN_PREFETCH=0 timeout 300 python scripts/trace_chain.py $B 4 1100 2>&1 | grep "per layer"; OCRB_CHAIN_MAX_B=0 timeout 300 python scripts/trace_chain.py $B 4 1100 2>&1 | grep "per layer"; done
