#!/bin/bash
# Round-2 first GPU pass: unit tests of the changed kernels (skinny BP=96/128, flash-decoding attention), the VLM tests,
# then the batched configs (P pages per step, B = 3P).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_dense.py -x -q -m gpu -k "skinny or decode_attention or argmax" > gpurun_out/r02a_dense.log 2>&1
echo "dense exit=$?"; tail -n 12 gpurun_out/r02a_dense.log
timeout 900 python -m pytest tests/test_gpu_vlm.py tests/test_gpu_read_path.py tests/test_gpu_folder.py -x -q -m gpu > gpurun_out/r02a_vlm.log 2>&1
echo "vlm exit=$?"; tail -n 12 gpurun_out/r02a_vlm.log
for P in 1 8 21 32; do
  timeout 600 python bench.py --pages $P --steps 2 --warmup 1 --no-cpu > gpurun_out/r02a_p$P.json 2> gpurun_out/r02a_p$P.err
  echo "P=$P exit=$?"; tail -c 300 gpurun_out/r02a_p$P.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02a_p$P.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","decode_tok_per_s","decode_phase_tok_per_s","phase_ms_per_step")}, d["roofline"]["frac"], d["roofline"]["decode_step_ms"], d["roofline"]["kernel_only"]["frac"], d["e2e"]["value"])
except Exception as e:
    print("no json", e)
PY
done
