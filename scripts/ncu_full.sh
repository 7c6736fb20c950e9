#!/bin/bash
# Full (--set full) captures of the hot kernels of the bench command; plain run first (profiling recipe).
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/plain_full.log 2>&1 || exit 1
# decode weight-streaming GEMM: skip the prefill/logits launches, take one decode step's worth of distinct shapes
timeout 600 ncu --set full --clock-control none --import-source on -k regex:skinny_gemm_kernel -s 230 -c 6 -o gpurun_out/prof_skinny \
  python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_skinny.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_attn_mma_kernel -s 56 -c 2 -o gpurun_out/prof_attn_mma \
  python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -s 300 -c 5 -o gpurun_out/prof_gemm \
  python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_gemm.log 2>&1
tail -c 300 gpurun_out/ncu_skinny.log
