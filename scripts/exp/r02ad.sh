#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_text_image.py -q -m gpu -x > gpurun_out/r02ad_img.log 2>&1; echo "image/text tests exit=$?"; tail -n 5 gpurun_out/r02ad_img.log
timeout 300 python scripts/bench_image.py --pages 64 --reps 5 --only high_contrast,binarize,sharpen,deskew
timeout 300 python scripts/bench_image.py --pages 1 --reps 9 --only high_contrast,binarize,sharpen,deskew
echo "bit-parallel:"; timeout 300 python scripts/bench_text.py
echo "wavefront only:"; OCRB_LEV_BITPAR=0 timeout 300 python scripts/bench_text.py
