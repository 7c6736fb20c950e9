#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -q -m gpu -x -k "chain or plan" > gpurun_out/r02t_chain.log 2>&1
echo "chain exit=$?"; tail -n 3 gpurun_out/r02t_chain.log
for PF in 0 12 24 40 64; do echo "PF=$PF"; OCRB_CHAIN_PREFETCH=$PF timeout 300 python scripts/trace_chain.py 3 4 1100 2>&1 | grep "per layer"; done
echo "keys=64"; OCRB_ATTN_KEYS=64 OCRB_CHAIN_PREFETCH=0 timeout 300 python scripts/trace_chain.py 3 4 1100 2>&1 | grep "per layer"
echo "keys=64 PF=24"; OCRB_ATTN_KEYS=64 OCRB_CHAIN_PREFETCH=24 timeout 300 python scripts/trace_chain.py 3 4 1100 > gpurun_out/r02t_trace.log 2>&1; sed -n 1,2p gpurun_out/r02t_trace.log; sed -n 9,15p gpurun_out/r02t_trace.log
