#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x -s > gpurun_out/r02i_all.log 2>&1
echo "all gpu tests exit=$?"; grep -E "^7B|passed|failed|error|Error" gpurun_out/r02i_all.log | tail -30
