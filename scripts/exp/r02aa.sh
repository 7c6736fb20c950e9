#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -q -m gpu -x -k "decode_attention or plan or chain" > gpurun_out/r02aa_attn.log 2>&1; echo "attention tests exit=$?"; tail -n 6 gpurun_out/r02aa_attn.log
timeout 900 python -m pytest tests/test_gpu_vlm.py tests/test_gpu_read_path.py -q -m gpu -x > gpurun_out/r02aa_vlm.log 2>&1; echo "vlm tests exit=$?"; tail -n 3 gpurun_out/r02aa_vlm.log
for CFG in 64 82 102; do for B in 3 96; do echo "attn cfg=$CFG B=$B"; OCRB_ATTN_CFG=$CFG timeout 300 python scripts/trace_chain.py $B 4 1100 2>&1 | grep "per layer"; done; done
echo "fused plan B=3"; OCRB_CHAIN_MAX_B=128 timeout 300 python scripts/trace_chain.py 3 4 1100 2>&1 | grep "per layer"
