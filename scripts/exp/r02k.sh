#!/bin/bash
mkdir -p gpurun_out
OCRB_SK_CLUSTER_DEBUG=1 timeout 1200 python -m pytest tests/test_gpu_dense.py -q -m gpu -x -k "skinny" -s > gpurun_out/r02k_dense.log 2>&1
echo "dense exit=$?"; grep "resident clusters" gpurun_out/r02k_dense.log | sort | uniq; tail -n 3 gpurun_out/r02k_dense.log
timeout 900 python -m pytest tests/test_gpu_text_image.py tests/test_gpu_vlm.py tests/test_gpu_read_path.py -x -q -m gpu > gpurun_out/r02k_img.log 2>&1
echo "img/vlm exit=$?"; tail -n 4 gpurun_out/r02k_img.log
timeout 600 python scripts/bench_skinny.py 3,24,96 > gpurun_out/r02k_skinny.log 2>&1; echo "skinny exit=$?"; grep "qkv\|o_proj\|down" gpurun_out/r02k_skinny.log
for P in 1 32; do
timeout 600 python bench.py --pages $P --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/r02k_p$P.json 2> gpurun_out/r02k_p$P.err
echo "P=$P exit=$?"; tail -c 300 gpurun_out/r02k_p$P.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02k_p$P.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","decode_tok_per_s","phase_ms_per_step")}, d["roofline"]["frac"], d["roofline"]["decode_step_ms"], d["roofline_tensor"]["frac"], d["e2e"]["value"], d["e2e"]["seconds"])
    print({k: (v.get("ms"), v.get("batch64_ms"), v.get("batch64_frac_hbm_peak")) for k, v in d["preprocess_kernels"].items()})
except Exception as e:
    print("no json", e)
PY
done
timeout 1500 python -m pytest tests/test_gpu_vlm_7b.py -q -m gpu -s > gpurun_out/r02k_7b.log 2>&1
echo "7b exit=$?"; grep -E "^7B|passed|failed" gpurun_out/r02k_7b.log
