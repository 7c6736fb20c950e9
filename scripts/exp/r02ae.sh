#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_text_image.py -q -m gpu -x -k "leven or compare or merge or text or tier" > gpurun_out/r02ae_text.log 2>&1; echo "text tests exit=$?"; tail -n 5 gpurun_out/r02ae_text.log
echo "bit-parallel:"; timeout 300 python scripts/bench_text.py
echo "wavefront only:"; OCRB_LEV_BITPAR=0 timeout 300 python scripts/bench_text.py
