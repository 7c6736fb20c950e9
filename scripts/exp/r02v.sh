#!/bin/bash
# 2-GPU: the TP test, then bench.py under torchrun at N=2 (configs[2] folder workload with a small page count + TP parity leg)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -4
timeout 900 python -m pytest tests/test_gpu_tp.py -x -q -m gpu > gpurun_out/r02v_tp.log 2>&1; echo "tp test exit=$?"; tail -n 4 gpurun_out/r02v_tp.log
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 > gpurun_out/r02v_bench_n2.json 2> gpurun_out/r02v_bench_n2.err
echo "bench N=2 exit=$? in ${SECONDS}s"; tail -c 600 gpurun_out/r02v_bench_n2.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02v_bench_n2.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","n_gpus","ms_per_step","decode_tok_per_s","phase_ms_per_step")}, d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["seconds"])
    print(d["config"]["workload"][:200]); print(d.get("extra"))
except Exception as e:
    print("no json", e)
PY
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r02v_ref_n2.json 2> gpurun_out/r02v_ref_n2.err
echo "ref N=2 exit=$? in ${SECONDS}s"; tail -c 300 gpurun_out/r02v_ref_n2.json
