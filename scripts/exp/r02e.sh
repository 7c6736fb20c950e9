#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/trace_skinny.py 96 qkv,o > gpurun_out/r02e_trace96.log 2>&1; echo "trace exit=$?"; cat gpurun_out/r02e_trace96.log
timeout 300 python scripts/trace_skinny.py 3 qkv,o > gpurun_out/r02e_trace3.log 2>&1; echo "trace exit=$?"; cat gpurun_out/r02e_trace3.log
