#!/bin/bash
# Launch list of the library's own kernels for the bench command (plain run first, as the profiling recipe asks).
# usage: scripts/ncu_launches.sh <count> <out.csv>
K='regex:^(skinny_gemm|decode_attn|argmax_step|decode_rope|rows_copy|gemm_tcgen05|flash_varlen|rmsnorm|rope_|kv_write|step_inc|normalize_pat|resize_|clahe|adaptive|sharpen|rgb2gray|dark_ext|deskew|warp_aff|levenshtein|lcs_align|gemv|skinny_norm|residual_add|allreduce)'
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c "${1:-1500}" --csv --log-file "${2:-gpurun_out/launches.csv}" \
  python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu1.log 2>&1
tail -c 200 gpurun_out/ncu1.log
