#!/bin/bash
# chain kernel: unit tests, engine tests, bench at P=1 / 8 / 32
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -q -m gpu -x -k "chain" > gpurun_out/r02o_chain.log 2>&1
echo "chain exit=$?"; tail -n 12 gpurun_out/r02o_chain.log
timeout 900 python -m pytest tests/test_gpu_vlm.py tests/test_gpu_read_path.py tests/test_gpu_folder.py -x -q -m gpu > gpurun_out/r02o_vlm.log 2>&1
echo "vlm exit=$?"; tail -n 8 gpurun_out/r02o_vlm.log
for P in 1 8 32; do
timeout 600 python bench.py --pages $P --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/r02o_p$P.json 2> gpurun_out/r02o_p$P.err
echo "P=$P exit=$?"; tail -c 400 gpurun_out/r02o_p$P.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02o_p$P.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","decode_tok_per_s","phase_ms_per_step")}, d["roofline"]["frac"], d["roofline"]["decode_step_ms"], d["e2e"]["value"])
except Exception as e:
    print("no json", e)
PY
done
