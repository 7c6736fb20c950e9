#!/bin/bash
# One GPU-box visit: smoke, GPU tests, bench (tiny plumbing check, then the real line). Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -n 5 gpurun_out/smoke.log
echo "== bench tiny"; timeout 600 python bench.py --tiny --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_tiny.log 2>&1; echo "bench tiny exit=$?"; tail -n 5 gpurun_out/bench_tiny.log
if [ "$1" != "notests" ]; then
echo "== tests"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/tests_gpu.log 2>&1; echo "tests exit=$?"; tail -n 15 gpurun_out/tests_gpu.log
fi
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -n 3 gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref exit=$?"; tail -n 2 gpurun_out/bench_ref.log
