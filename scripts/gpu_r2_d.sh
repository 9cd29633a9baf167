#!/bin/bash
set -u
mkdir -p gpurun_out
for o in "cg_single_reduction=1" "cg_single_reduction=1 cg_zigzag=0" "cg_single_reduction=1 cg_dynamic=0 cg_zigzag=0"; do
  timeout 300 python scripts/prof_slab.py 1024 8192 $o 2>&1 | grep -E "^rank|rror" | tee -a gpurun_out/r2d_prof_slab.txt
done
