#!/bin/bash
# window attention test + vision timing; big-B skinny ring variants (rebuilt on the box)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py -q -m gpu -x -k "attention" > gpurun_out/r02w_attn.log 2>&1; echo "attention tests exit=$?"; tail -n 3 gpurun_out/r02w_attn.log
timeout 600 python -m pytest tests/test_gpu_vlm.py -q -m gpu -x > gpurun_out/r02w_vlm.log 2>&1; echo "vlm tests exit=$?"; tail -n 3 gpurun_out/r02w_vlm.log
echo "== (ST_BIG, OCC_BIG) = (6, 1)"
for B in 96 128; do OCRB_CHAIN_MAX_B=0 timeout 300 python scripts/trace_chain.py $B 4 1100 2>&1 | grep "per layer"; done
cd handwritten-ocr_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-ffp-contract=off -DSK_ST_BIG=3 -DSK_OCC_BIG=2 -c skinny.cu -o skinny.o 2> /tmp/sk.log || tail /tmp/sk.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libocrb200.so lib.o textops.o denoise.o image.o resize_patchify.o inpaint.o dense.o gemm_tcgen05.o attention.o flash_tc.o decode.o skinny.o chain.o comm.o
cd ../..
echo "== (ST_BIG, OCC_BIG) = (3, 2)"
for B in 96 128; do OCRB_CHAIN_MAX_B=0 timeout 300 python scripts/trace_chain.py $B 4 1100 2>&1 | grep "per layer"; done
cd handwritten-ocr_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-ffp-contract=off -DSK_ST_BIG=6 -DSK_OCC_BIG=1 -c skinny.cu -o skinny.o 2> /tmp/sk.log || tail /tmp/sk.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libocrb200.so lib.o textops.o denoise.o image.o resize_patchify.o inpaint.o dense.o gemm_tcgen05.o attention.o flash_tc.o decode.o skinny.o chain.o comm.o
cd ../..
for P in 1 32; do
timeout 600 python bench.py --pages $P --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/r02w_p$P.json 2> gpurun_out/r02w_p$P.err
echo "P=$P exit=$?"; tail -c 300 gpurun_out/r02w_p$P.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02w_p$P.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","decode_tok_per_s","phase_ms_per_step")}, d["roofline"]["frac"], d["roofline"]["decode_step_ms"], d["roofline_tensor"]["frac"], d["roofline_tensor"]["vision"], d["e2e"]["value"])
except Exception as e:
    print("no json", e)
PY
done
