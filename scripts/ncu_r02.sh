#!/bin/bash
# Round-2 ncu captures (B200_PROFILING.md): plain run first, then the launch list (gpu__time_duration, cold-cache and
# serialised: compare SHARES) and `--set full` captures of the kernels of one vision block / prefill layer / decode layer.
# gpurun brings back at most 64 MiB: every report is exported to CSV (`--page raw`) on the box and only the decode-layer
# report (the dominant kernels, with source) is kept as .ncu-rep.  Summaries derived from it: profiles/r02*.
mkdir -p gpurun_out
T1="python scripts/ncu_target.py --pages 1 --new-tokens 4"
T32="python scripts/ncu_target.py --pages 32 --new-tokens 3"
NV='--nvtx --nvtx-include capture/'
$T1 > gpurun_out/r02n_plain_p1.log 2>&1 || { echo "plain P=1 failed"; tail -n 5 gpurun_out/r02n_plain_p1.log; exit 1; }
tail -n 1 gpurun_out/r02n_plain_p1.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none $NV --csv --log-file gpurun_out/r02n_launches_p1.csv $T1 > gpurun_out/r02n_ncu_l1.log 2>&1
echo "launch list P=1: $?"
full() {  # name kernel-regex skip count keep-rep command...
  local name=$1 kr=$2 skip=$3 cnt=$4 keep=$5; shift 5
  local src=""; [ "$keep" = 1 ] && src="--import-source on"
  timeout 600 ncu --set full --clock-control none $src $NV -k "regex:$kr" -s "$skip" -c "$cnt" -f -o "gpurun_out/r02n_$name" "$@" > "gpurun_out/r02n_ncu_$name.log" 2>&1
  echo "full $name: $?"
  ncu -i "gpurun_out/r02n_$name.ncu-rep" --page raw --csv > "gpurun_out/r02n_$name.raw.csv" 2>/dev/null
  [ "$keep" = 1 ] || rm -f "gpurun_out/r02n_$name.ncu-rep"
}
# filtered launch order before decode: patch-embed GEMM, then per vision block qkv / attention / proj / gate-up / down (5),
# 2 merger GEMMs, then per prefill layer qkv / flash_tc / o / gate-up / down (5)
PRE='(flash_tc_kernel|flash_varlen_kernel|gemm_tcgen05_kernel)'
full vision $PRE 31 10 0 $T1          # vision blocks 6 (windowed) and 7 (full attention)
full prefill $PRE 163 5 0 $T1         # prefill layer 0
# filtered launch order in a decode step: per layer qkv (cluster) / attention / o_proj (cluster) / gate-up (stream-K) / down (cluster)
DEC='(skinny_gemm_kernel|skinny_cluster_kernel|decode_attn_kernel)'
full decode_b3 $DEC 135 6 1 $T1       # layer 27 + lm_head
$T32 > gpurun_out/r02n_plain_p32.log 2>&1 || { echo "plain P=32 failed"; tail -n 5 gpurun_out/r02n_plain_p32.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none $NV -k 'regex:(skinny_|decode_attn|argmax|embed|rope_table|gemm_tcgen05|flash_|rmsnorm|rope_|kv_write|rows_copy)' --csv --log-file gpurun_out/r02n_launches_p32.csv $T32 > gpurun_out/r02n_ncu_l32.log 2>&1
echo "launch list P=32: $?"
full decode_b96 $DEC 135 6 0 $T32
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
du -sh gpurun_out; ls -la gpurun_out | head -40
