#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -q -m gpu -x -k "chain or plan" > gpurun_out/r02r_chain.log 2>&1
echo "chain exit=$?"; tail -n 12 gpurun_out/r02r_chain.log
timeout 900 python -m pytest tests/test_gpu_vlm.py -x -q -m gpu > gpurun_out/r02r_vlm.log 2>&1
echo "vlm exit=$?"; tail -n 8 gpurun_out/r02r_vlm.log
timeout 300 python scripts/trace_chain.py 3 4 1100 > gpurun_out/r02r_trace_b3.log 2>&1; echo "trace exit=$?"; head -n 32 gpurun_out/r02r_trace_b3.log
OCRB_CHAIN_PREFETCH=0 timeout 300 python scripts/trace_chain.py 3 4 1100 2>&1 | grep "per layer"
OCRB_CHAIN_PREFETCH=48 timeout 300 python scripts/trace_chain.py 3 4 1100 2>&1 | grep "per layer"
OCRB_CHAIN_ATTN=0 timeout 300 python scripts/trace_chain.py 3 4 1100 2>&1 | grep "per layer"
OCRB_CHAIN_MAX_B=0 timeout 300 python scripts/trace_chain.py 3 4 1100 2>&1 | grep "per layer"
for P in 1 8; do
timeout 600 python bench.py --pages $P --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/r02r_p$P.json 2> gpurun_out/r02r_p$P.err
echo "P=$P exit=$?"; tail -c 400 gpurun_out/r02r_p$P.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02r_p$P.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","decode_tok_per_s","phase_ms_per_step")}, d["roofline"]["frac"], d["roofline"]["decode_step_ms"], d["e2e"]["value"])
except Exception as e:
    print("no json", e)
PY
done
