#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dense.py -q -m gpu -x -k "skinny" > gpurun_out/r02j_dense.log 2>&1
echo "dense exit=$?"; tail -n 8 gpurun_out/r02j_dense.log
timeout 600 python scripts/bench_skinny.py 3,24,96 > gpurun_out/r02j_skinny.log 2>&1; echo "skinny exit=$?"; cat gpurun_out/r02j_skinny.log
timeout 900 python -m pytest tests/test_gpu_vlm.py tests/test_gpu_read_path.py -x -q -m gpu > gpurun_out/r02j_vlm.log 2>&1
echo "vlm exit=$?"; tail -n 5 gpurun_out/r02j_vlm.log
for P in 1 32; do
timeout 600 python bench.py --pages $P --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/r02j_p$P.json 2> gpurun_out/r02j_p$P.err
echo "P=$P exit=$?"; tail -c 300 gpurun_out/r02j_p$P.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02j_p$P.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","decode_tok_per_s","phase_ms_per_step")}, d["roofline"]["frac"], d["roofline"]["decode_step_ms"], d["roofline_tensor"]["frac"], d["roofline_tensor"]["vision"]["ms_per_read"], d["roofline_tensor"]["prefill"]["ms_per_read"], d["e2e"]["value"], d["e2e"]["seconds"])
except Exception as e:
    print("no json", e)
PY
done
OCRB_VISION_CHUNK=3 timeout 600 python bench.py --pages 32 --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/r02j_p32v3.json 2> gpurun_out/r02j_p32v3.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02j_p32v3.json").read().strip().splitlines()[-1])
print("vision_chunk=3:", d["value"], d["roofline_tensor"]["vision"]["ms_per_read"], d["roofline_tensor"]["prefill"]["ms_per_read"])
PY
