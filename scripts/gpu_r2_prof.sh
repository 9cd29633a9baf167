#!/bin/bash
# round 2: ncu captures of every kernel a bench line names (one GPU).  Summaries (text) come back; the reports stay on the box
# except the 4096^2 one.
set -u
mkdir -p gpurun_out/prof
export PANO_BENCH_MIN_WARMUP=3
NCU="ncu --set full --clock-control none -f"
B="python bench.py --steps 2 --no-cpu --no-extra"
# 4096^2: advect, -div, CG, project of one step (step 5)
timeout 900 $NCU --import-source on -k regex:'k_advect|k_neg_divergence|k_project|k_cg' -s 16 -c 4 -o gpurun_out/prof/r02_4096 $B --grid 4096 --warmup 3 > gpurun_out/prof/ncu_4096.log 2>&1; echo "4096 rc=$?"
timeout 900 $NCU -k regex:k_cg -s 4 -c 1 -o gpurun_out/prof/r02_cg_8192 $B --grid 8192 --warmup 3 > gpurun_out/prof/ncu_8192.log 2>&1; echo "8192 rc=$?"
timeout 600 $NCU -k regex:k_cg -s 4 -c 1 -o gpurun_out/prof/r02_cg_1024 $B --grid 1024 --warmup 3 > gpurun_out/prof/ncu_1024.log 2>&1; echo "1024 rc=$?"
timeout 600 $NCU -k regex:k_cg -s 4 -c 1 -o gpurun_out/prof/r02_cg_128 $B --grid 128 --warmup 3 > gpurun_out/prof/ncu_128.log 2>&1; echo "128 rc=$?"
# the per-GPU slab of 8192^2 on 8 GPUs (1024 x 8192), through pano_dist with one rank: slab forms of all four kernels (step 12 of time_slab.py)
timeout 900 $NCU -k regex:'k_advect|k_neg_divergence|k_project|k_cg' -s 44 -c 4 -o gpurun_out/prof/r02_slab_1024x8192 python scripts/time_slab.py 1024 8192 > gpurun_out/prof/ncu_slab.log 2>&1; echo "slab rc=$?"
# the single-reduction streaming kernel on one GPU's 4096^2 for comparison (the auto choice there is the two-reduction kernel)
timeout 900 $NCU -k regex:k_cg -s 4 -c 1 -o gpurun_out/prof/r02_cg_sr_4096 $B --grid 4096 --warmup 3 --opt cg_single_reduction=1 > gpurun_out/prof/ncu_4096_sr.log 2>&1; echo "4096 single-reduction rc=$?"
for r in gpurun_out/prof/*.ncu-rep; do
  python scripts/ncu_summary.py $r > ${r%.ncu-rep}_ncu.txt 2>&1
done
ls -la gpurun_out/prof/
# keep only the 4096^2 report (source view of the advection and CG kernels); the others are summarised above
find gpurun_out/prof -name '*.ncu-rep' ! -name 'r02_4096.ncu-rep' -delete
# launch list of the default bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
unset PANO_BENCH_MIN_WARMUP
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof/r02_launches_8192.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/prof/launches.log 2>&1; echo "launch list rc=$?"
du -sh gpurun_out/prof
