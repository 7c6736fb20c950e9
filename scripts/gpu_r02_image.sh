#!/bin/bash
# Round-2 image-kernel visit: parity tests of the image / read path, timings at 1 and 64 pages, ncu metric pass at 64 pages.
mkdir -p gpurun_out
T=${1:-r02img}
timeout 600 python -m pytest tests/test_gpu_text_image.py tests/test_denoise.py tests/test_remove_lines.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1
echo "tests exit=$?"; tail -n 12 gpurun_out/${T}_tests.log
timeout 300 python scripts/bench_image.py --pages 1 --reps 21 --only high_contrast,binarize,sharpen,deskew > gpurun_out/${T}_p1.json 2>gpurun_out/${T}_p1.err; cat gpurun_out/${T}_p1.json
timeout 300 python scripts/bench_image.py --pages 64 --reps 11 --only high_contrast,binarize,sharpen,deskew > gpurun_out/${T}_p64.json 2>gpurun_out/${T}_p64.err; cat gpurun_out/${T}_p64.json
M='gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active'
CMD3="python scripts/bench_image.py --pages 64 --reps 1 --only high_contrast,binarize,sharpen,deskew"
timeout 400 ncu --metrics "$M" --clock-control none -k 'regex:(rgb2gray|clahe|adaptive|sharpen|dark_ext|deskew|warp_aff)' --csv --log-file gpurun_out/${T}_image64_ncu.csv $CMD3 > gpurun_out/${T}_ncu_img.log 2>&1
echo "ncu image: $?"
