#!/bin/bash
mkdir -p gpurun_out
echo "== wide tiles"; timeout 300 python scripts/bench_gemm.py 2>&1 | grep "vis "
echo "== 128-wide tiles for N <= 1280"; OCRB_GEMM_NARROW_N=1280 timeout 300 python scripts/bench_gemm.py 2>&1 | grep "vis "
