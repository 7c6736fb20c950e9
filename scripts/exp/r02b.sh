#!/bin/bash
# Round-2 GPU pass b: unit tests of the changed kernels, VLM / read-path / folder tests, per-kernel decode timings.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -x -q -m gpu -k "skinny or decode_attention or argmax" > gpurun_out/r02b_dense.log 2>&1
echo "dense exit=$?"; tail -n 12 gpurun_out/r02b_dense.log
timeout 900 python -m pytest tests/test_gpu_vlm.py tests/test_gpu_read_path.py tests/test_gpu_folder.py -x -q -m gpu > gpurun_out/r02b_vlm.log 2>&1
echo "vlm exit=$?"; tail -n 15 gpurun_out/r02b_vlm.log
timeout 600 python scripts/bench_attn.py > gpurun_out/r02b_attn.log 2>&1; echo "attn exit=$?"; cat gpurun_out/r02b_attn.log
timeout 600 python scripts/bench_skinny.py 3,24,63,96 > gpurun_out/r02b_skinny.log 2>&1; echo "skinny exit=$?"; cat gpurun_out/r02b_skinny.log
