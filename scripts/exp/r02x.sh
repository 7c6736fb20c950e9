#!/bin/bash
mkdir -p gpurun_out
for CFG in 64 82; do for B in 3 24 96; do echo "attn cfg=$CFG B=$B"; OCRB_ATTN_CFG=$CFG OCRB_CHAIN_MAX_B=0 timeout 300 python scripts/trace_chain.py $B 4 1100 2>&1 | grep "per layer"; done; done
for VC in 2 3 4 8; do
OCRB_VISION_CHUNK=$VC timeout 600 python bench.py --pages 32 --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/r02x_vc$VC.json 2> gpurun_out/r02x_vc$VC.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02x_vc$VC.json").read().strip().splitlines()[-1])
    print("vision_chunk=$VC", d["value"], d["phase_ms_per_step"], d["roofline_tensor"]["vision"]["ms_per_read"], d["roofline_tensor"]["prefill"]["ms_per_read"])
except Exception as e:
    print("no json", e)
PY
done
