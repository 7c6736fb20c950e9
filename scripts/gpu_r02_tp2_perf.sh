#!/bin/bash
# 2-GPU A/B of the tensor-parallel decode step (72B-class shards at TP-2): all-reduce fused into the GEMM epilogue / separate
# kernel x pair-exchange arg max / logits all-gather.
mkdir -p gpurun_out
T=${1:-r02tpp}
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/bench_tp.py --new-tokens 128"
run() { name=$1; shift; env "$@" timeout 300 $R > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err; echo "$name exit=$?"; tail -n 1 gpurun_out/${T}_$name.json | cut -c1-260; }
run fused_pair OCRB_TP_FUSED=1
run unfused_pair OCRB_TP_FUSED=0
run fused_allfence_pair OCRB_TP_FUSED=1 OCRB_TP_FUSED_MODE=1
run unfused_gather OCRB_TP_FUSED=0 OCRB_TP_ARGMAX_GATHER=1
