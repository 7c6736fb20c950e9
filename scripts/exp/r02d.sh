#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dense.py -q -m gpu > gpurun_out/r02d_dense.log 2>&1
echo "dense exit=$?"; tail -n 25 gpurun_out/r02d_dense.log
timeout 300 python scripts/trace_skinny.py 96 o,qkv > gpurun_out/r02d_trace96.log 2>&1; echo "trace exit=$?"; cat gpurun_out/r02d_trace96.log
timeout 600 python scripts/bench_skinny.py 3,96 > gpurun_out/r02d_skinny.log 2>&1; echo "skinny exit=$?"; cat gpurun_out/r02d_skinny.log
timeout 600 python scripts/bench_gemm.py > gpurun_out/r02d_gemm.log 2>&1; echo "gemm exit=$?"; cat gpurun_out/r02d_gemm.log
