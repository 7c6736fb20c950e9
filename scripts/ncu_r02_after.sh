#!/bin/bash
# "after" captures of the kernels changed in round 2 (eight GEMM epilogue warps, persistent window attention, transposed
# decode attention with vectorised mRoPE staging): same target and filters as scripts/ncu_r02.sh
mkdir -p gpurun_out
T1="python scripts/ncu_target.py --pages 1 --new-tokens 4"
T32="python scripts/ncu_target.py --pages 32 --new-tokens 3"
NV='--nvtx --nvtx-include capture/'
$T1 > gpurun_out/r02q_plain_p1.log 2>&1 || { echo "plain P=1 failed"; tail -n 5 gpurun_out/r02q_plain_p1.log; exit 1; }
full() {
  local name=$1 kr=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none $NV -k "regex:$kr" -s "$skip" -c "$cnt" -f -o "gpurun_out/r02q_$name" "$@" > "gpurun_out/r02q_ncu_$name.log" 2>&1
  echo "full $name: $?"
  ncu -i "gpurun_out/r02q_$name.ncu-rep" --page raw --csv > "gpurun_out/r02q_$name.raw.csv" 2>/dev/null
  rm -f "gpurun_out/r02q_$name.ncu-rep"
}
full vision '(flash_tc_kernel|window_attn_kernel|flash_varlen_kernel|gemm_tcgen05_kernel)' 31 5 $T1     # vision block 6: qkv, window attention, proj, gate/up, down
full dattn_b3 decode_attn_kernel 10 1 $T1
full dattn_b96 decode_attn_kernel 10 1 $T32
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none $NV --csv --log-file gpurun_out/r02q_launches_p1.csv $T1 > gpurun_out/r02q_ncu_l1.log 2>&1
echo "launch list P=1: $?"
du -sh gpurun_out
