#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_text_image.py -q -m gpu -x -k "compare or merge or text or leven" > gpurun_out/r02z_text.log 2>&1; echo "text tests exit=$?"; tail -n 3 gpurun_out/r02z_text.log
for V in "OCRB_SK_CLUSTER=1" "OCRB_SK_CLUSTER=0" "OCRB_CHAIN_MAX_B=128 OCRB_CHAIN_ATTN=1"; do
  env $V timeout 300 python scripts/bench_tp_rank.py 3 2>&1 | tail -n 1
done
timeout 600 python bench.py --pages 32 --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/r02z_p32.json 2> gpurun_out/r02z_p32.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02z_p32.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","decode_tok_per_s","phase_ms_per_step")}, d["roofline"]["frac"], d["roofline"]["decode_step_ms"], d["e2e"]["value"])
except Exception as e:
    print("no json", e)
PY
