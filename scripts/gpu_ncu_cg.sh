#!/bin/bash
# ncu --set full on the persistent CG kernel (one launch each at 4096^2 and 1024^2)
set -u
mkdir -p gpurun_out
N=${1:-4096}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cg -s 2 -c 1 -f -o gpurun_out/prof_cg_$N \
    python bench.py --steps 1 --warmup 3 --n $N --no-cpu > gpurun_out/ncu_cg_$N.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_cg_$N.log; ls -la gpurun_out/*.ncu-rep
