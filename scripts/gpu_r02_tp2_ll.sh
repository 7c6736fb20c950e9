#!/bin/bash
# 2-GPU visit: TP test (three all-reduce routes bit-equal), then the decode step at TP-2 on 72B-class shards per route.
mkdir -p gpurun_out
T=${1:-r02ll}
timeout 400 python -m pytest tests/test_gpu_tp.py -x -q -s -m gpu > gpurun_out/${T}_tp.log 2>&1; rc=$?; echo "tp test exit=$rc"; tail -n 6 gpurun_out/${T}_tp.log
[ $rc -ne 0 ] && { grep -n "Error\|error\|assert\|never" gpurun_out/${T}_tp.log | head -20; exit 1; }
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/bench_tp.py --new-tokens 128"
run() { name=$1; shift; env OCRB_X=1 "$@" timeout 300 $R > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err; echo "$name exit=$?"; tail -n 1 gpurun_out/${T}_$name.json | cut -c1-200; }
run ll_default
run unfused OCRB_TP_FUSED=0
run pull OCRB_TP_FUSED=1
