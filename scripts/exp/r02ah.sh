#!/bin/bash
for K in 128 64 32; do for B in 3 24 96; do echo "keys=$K B=$B"; OCRB_ATTN_KEYS=$K timeout 300 python scripts/trace_chain.py $B 4 1100 2>&1 | grep "per layer"; done; done
