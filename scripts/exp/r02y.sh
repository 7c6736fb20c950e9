#!/bin/bash
# 8-GPU: bench.py under torchrun at N=8 (configs[2] folder workload, TP-2 parity leg, 72B-class TP-8 leg)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 1 --warmup 1 > gpurun_out/r02y_bench_n8.json 2> gpurun_out/r02y_bench_n8.err
echo "bench N=8 exit=$? in ${SECONDS}s"; tail -c 800 gpurun_out/r02y_bench_n8.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02y_bench_n8.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","n_gpus","ms_per_step","decode_tok_per_s","phase_ms_per_step")}, d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["seconds"])
    print(json.dumps(d.get("extra"), indent=1))
except Exception as e:
    print("no json", e)
PY
