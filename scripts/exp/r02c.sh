#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -x -q -m gpu -k "skinny" > gpurun_out/r02c_dense.log 2>&1
echo "dense exit=$?"; tail -n 5 gpurun_out/r02c_dense.log
timeout 900 python -m pytest tests/test_gpu_vlm.py tests/test_gpu_read_path.py tests/test_gpu_folder.py -x -q -m gpu > gpurun_out/r02c_vlm.log 2>&1
echo "vlm exit=$?"; tail -n 15 gpurun_out/r02c_vlm.log
timeout 600 python scripts/bench_skinny.py 3,24,96 > gpurun_out/r02c_skinny.log 2>&1; echo "skinny exit=$?"; cat gpurun_out/r02c_skinny.log
OCRB_SK_GRID=112 timeout 600 python scripts/bench_skinny.py 3,24,96 > gpurun_out/r02c_skinny112.log 2>&1; echo "skinny112 exit=$?"; grep -v "gate_up\|lm_head" gpurun_out/r02c_skinny112.log
OCRB_SK_GRID=144 timeout 600 python scripts/bench_skinny.py 3,24,96 > gpurun_out/r02c_skinny144.log 2>&1; echo "skinny144 exit=$?"; grep "qkv" gpurun_out/r02c_skinny144.log
for P in 1 32; do
timeout 600 python bench.py --pages $P --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/r02c_p$P.json 2> gpurun_out/r02c_p$P.err
echo "P=$P exit=$?"; tail -c 300 gpurun_out/r02c_p$P.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c_p$P.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","decode_tok_per_s","phase_ms_per_step")}, d["roofline"]["frac"], d["roofline"]["decode_step_ms"], d["roofline_tensor"]["frac"], d["e2e"])
except Exception as e:
    print("no json", e)
PY
done
