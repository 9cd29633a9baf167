#!/bin/bash
# round 2: first look at the new kernels under ncu (4096^2): advect_tma, cg_sr vs cg_stream
set -u
mkdir -p gpurun_out/prof
export PANO_BENCH_MIN_WARMUP=3
NCU="ncu --set full --clock-control none -f"
B="python bench.py --steps 2 --no-cpu --no-extra"
timeout 900 $NCU --import-source on -k regex:'k_advect|k_cg' -s 8 -c 2 -o gpurun_out/prof/r02a_4096 $B --grid 4096 --warmup 3 > gpurun_out/prof/ncu_a_4096.log 2>&1; echo "4096 rc=$?"
timeout 900 $NCU --import-source on -k regex:k_cg -s 4 -c 1 -o gpurun_out/prof/r02a_cg_stream2_4096 $B --grid 4096 --warmup 3 --opt cg_single_reduction=0 > gpurun_out/prof/ncu_a_4096_two.log 2>&1; echo "4096 two-reduction rc=$?"
timeout 900 $NCU -k regex:k_advect -s 4 -c 1 -o gpurun_out/prof/r02a_advect3_4096 $B --grid 4096 --warmup 3 --opt advect_kernel=3 > gpurun_out/prof/ncu_a_adv3.log 2>&1; echo "march3 rc=$?"
for r in gpurun_out/prof/r02a_*.ncu-rep; do
  python scripts/ncu_summary.py $r > ${r%.ncu-rep}_ncu.txt 2>&1
done
ls -la gpurun_out/prof/
