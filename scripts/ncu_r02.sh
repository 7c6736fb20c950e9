#!/bin/bash
# Round-2 ncu captures (B200_PROFILING.md): plain run first, then the launch list (gpu__time_duration, cold-cache and
# serialised: compare SHARES) and `--set full` captures of the top kernels.  Everything lands in gpurun_out/; the
# summaries derived from it are committed under profiles/r02*.
mkdir -p gpurun_out
K='regex:(skinny_|decode_attn|argmax_step|decode_rope|rows_copy|gemm_tcgen05|flash_|rmsnorm|rope_|kv_write|step_inc|normalize_pat|resize_|clahe|adaptive|sharpen|rgb2gray|dark_ext|deskew|warp_aff|levenshtein|lcs_align|embed_gather|residual_add)'
CMD1="python bench.py --steps 1 --warmup 1 --no-cpu --no-extra"
CMD2="python bench.py --pages 32 --steps 1 --warmup 1 --no-cpu --no-extra"
$CMD1 > gpurun_out/r02n_plain_p1.log 2>&1 || { echo "plain P=1 failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 2600 --csv --log-file gpurun_out/r02n_launches_p1.csv $CMD1 > gpurun_out/r02n_ncu_l1.log 2>&1
echo "launch list P=1: $?"
full() {  # name kernel-regex skip count command...
  local name=$1 kr=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$kr" -s "$skip" -c "$cnt" -f -o "gpurun_out/r02n_$name" "$@" > "gpurun_out/r02n_ncu_$name.log" 2>&1
  echo "full $name: $?"
}
full flash_tc flash_tc_kernel 0 4 $CMD1
full flash_win flash_varlen_kernel 2 2 $CMD1
full gemm gemm_tcgen05_kernel 20 6 $CMD1
full dattn_b3 decode_attn_kernel 56 2 $CMD1
full skinny_b3 skinny_gemm_kernel 120 4 $CMD1
full cluster_b3 skinny_cluster_kernel 120 4 $CMD1
$CMD2 > gpurun_out/r02n_plain_p32.log 2>&1 || { echo "plain P=32 failed"; exit 1; }
full dattn_b96 decode_attn_kernel 56 2 $CMD2
full skinny_b96 skinny_gemm_kernel 120 4 $CMD2
full cluster_b96 skinny_cluster_kernel 120 4 $CMD2
ls -la gpurun_out/*.ncu-rep
# preprocessing kernels on 64 pages (151 MB in: larger than L2) and the text kernels: DRAM bytes / throughput per launch
M='gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active'
CMD3="python scripts/bench_image.py --pages 64 --reps 1 --only high_contrast,binarize,sharpen,deskew"
$CMD3 > gpurun_out/r02n_plain_img.log 2>&1 && \
timeout 600 ncu --metrics "$M" --clock-control none -k 'regex:(rgb2gray|clahe|adaptive|sharpen|dark_ext|deskew|warp_aff)' --csv --log-file gpurun_out/r02n_image64.csv $CMD3 > gpurun_out/r02n_ncu_img.log 2>&1
echo "image kernels: $?"
CMD4="python scripts/bench_text.py"
$CMD4 > gpurun_out/r02n_plain_text.log 2>&1 && \
timeout 600 ncu --metrics "$M" --clock-control none -k 'regex:(levenshtein|lcs_align)' --csv --log-file gpurun_out/r02n_text.csv $CMD4 > gpurun_out/r02n_ncu_text.log 2>&1
echo "text kernels: $?"
