#!/bin/bash
# GEMM epilogue warps: 8 (built) vs 4 / 16 (rebuilt on the box), then the GEMM tests and a bench line
mkdir -p gpurun_out
echo "== EW=8"; timeout 300 python scripts/bench_gemm.py 2>&1 | tee gpurun_out/r02u_gemm_ew8.log
timeout 600 python -m pytest tests/test_gpu_dense.py -q -m gpu -x -k "gemm" > gpurun_out/r02u_gemm_tests.log 2>&1; echo "gemm tests exit=$?"; tail -n 3 gpurun_out/r02u_gemm_tests.log
cd handwritten-ocr_b200/csrc
for EW in 16 4; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-ffp-contract=off -Xptxas -v -DGM_EPI_WARPS_N=$EW -c gemm_tcgen05.cu -o gemm_tcgen05.o 2> /tmp/ew.log
  grep -E "registers|spill" /tmp/ew.log | sort | uniq -c
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libocrb200.so lib.o textops.o denoise.o image.o resize_patchify.o inpaint.o dense.o gemm_tcgen05.o attention.o flash_tc.o decode.o skinny.o chain.o comm.o
  echo "== EW=$EW"; (cd ../.. && timeout 300 python scripts/bench_gemm.py 2>&1 | tee gpurun_out/r02u_gemm_ew$EW.log)
done
