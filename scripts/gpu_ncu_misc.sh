#!/bin/bash
# ncu --set full on the non-solver kernels at N (default 4096)
set -u
mkdir -p gpurun_out
N=${1:-4096}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_advect|k_neg_divergence|k_project' -s 6 -c 3 -f -o gpurun_out/prof_misc_$N \
    python bench.py --steps 1 --warmup 3 --n $N --no-cpu > gpurun_out/ncu_misc_$N.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_misc_$N.ncu-rep
