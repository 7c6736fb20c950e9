#!/bin/bash
# 2-GPU visit: TP-2 test (all-reduce kernel, pair-exchange arg max, tiny-config parity) and the N = 2 bench line (short).
mkdir -p gpurun_out
T=${1:-r02tp}
timeout 500 python -m pytest tests/test_gpu_tp.py -x -q -s -m gpu > gpurun_out/${T}_tp.log 2>&1; echo "tp test exit=$?"; tail -n 8 gpurun_out/${T}_tp.log
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; echo "bench n2 exit=$?"
tail -c 1500 gpurun_out/${T}_bench_n2.json; tail -n 5 gpurun_out/${T}_bench_n2.err
