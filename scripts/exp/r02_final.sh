#!/bin/bash
# final validation of the round: smoke, whole GPU suite, default bench (extras + cpu leg), reference arm
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -n 2 gpurun_out/smoke.log
echo "== tests"; SECONDS=0; timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/tests_gpu.log 2>&1; echo "tests exit=$? in ${SECONDS}s"; tail -n 4 gpurun_out/tests_gpu.log
echo "== bench"; SECONDS=0; timeout 1500 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit=$? in ${SECONDS}s"; tail -n 3 gpurun_out/bench.err
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref exit=$?"
