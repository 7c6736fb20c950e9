#!/bin/bash
# Round-2 baseline: the batched configs (P pages per step, B = 3P) on the round-1 kernels.
mkdir -p gpurun_out
for P in 8 21 32; do
  timeout 600 python bench.py --pages $P --steps 2 --warmup 1 --no-cpu > gpurun_out/r02a_p$P.json 2> gpurun_out/r02a_p$P.err
  echo "P=$P exit=$?"; tail -c 600 gpurun_out/r02a_p$P.json
done
