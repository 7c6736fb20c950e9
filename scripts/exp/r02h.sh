#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py -q -m gpu -x -k "decode_attention" > gpurun_out/r02h_dense.log 2>&1
echo "dense exit=$?"; tail -n 3 gpurun_out/r02h_dense.log
for K in 64 128 256; do
KEYS=$K timeout 300 python scripts/bench_attn.py > gpurun_out/r02h_attn64_$K.log 2>&1; echo "attn cfg64 keys $K exit=$?"; cat gpurun_out/r02h_attn64_$K.log
done
OCRB_ATTN_CFG=82 KEYS=64,128 timeout 300 python scripts/bench_attn.py > gpurun_out/r02h_attn82.log 2>&1; echo "attn cfg82 exit=$?"; cat gpurun_out/r02h_attn82.log
