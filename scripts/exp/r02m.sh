#!/bin/bash
# HEAD validation on a fresh box: smoke, whole GPU suite, default bench (extras + cpu leg), reference arm.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -n 3 gpurun_out/smoke.log
echo "== tests"; SECONDS=0; timeout 2400 python -m pytest tests -x -q -m gpu --durations=15 > gpurun_out/tests_gpu.log 2>&1; echo "tests exit=$? in ${SECONDS}s"; tail -n 30 gpurun_out/tests_gpu.log
echo "== bench"; SECONDS=0; timeout 1500 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit=$? in ${SECONDS}s"; tail -c 1500 gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err
echo "== bench reference"; SECONDS=0; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref exit=$? in ${SECONDS}s"; tail -c 600 gpurun_out/bench_ref.log
