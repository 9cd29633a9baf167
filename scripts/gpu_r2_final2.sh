#!/bin/bash
# final single-GPU records of round 2 (second session): full GPU test suite, smoke, the default bench line, ncu of the wide-block -div /
# projection kernels at 4096^2 and of the Grid3d kernels at 256^3, launch list of the default bench command
set -u
mkdir -p gpurun_out/final2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final2/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final2/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final2/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final2/smoke.log
timeout 900 python bench.py > gpurun_out/final2/bench_n1.json 2> gpurun_out/final2/bench_n1.err; echo "bench rc=$?"
export PANO_BENCH_MIN_WARMUP=3
timeout 900 ncu --set full --clock-control none -f -k regex:'k_neg_divergence|k_project' -s 8 -c 2 -o gpurun_out/final2/r02_negdiv_project_4096 python bench.py --steps 2 --no-cpu --no-extra --grid 4096 --warmup 3 > gpurun_out/final2/ncu_np.log 2>&1; echo "ncu 4096 rc=$?"
python scripts/ncu_summary.py gpurun_out/final2/r02_negdiv_project_4096.ncu-rep > gpurun_out/final2/r02_negdiv_project_4096_ncu.txt 2>&1
rm -f gpurun_out/final2/r02_negdiv_project_4096.ncu-rep
unset PANO_BENCH_MIN_WARMUP
timeout 900 ncu --set full --clock-control none -f -k regex:'k3_advect|k3_neg_div|k3_cg|k3_project' -s 32 -c 4 -o gpurun_out/final2/r02_grid3_256 python scripts/bench_grid3.py 256 2 > gpurun_out/final2/ncu_g3.log 2>&1; echo "ncu grid3 rc=$?"
python scripts/ncu_summary.py gpurun_out/final2/r02_grid3_256.ncu-rep > gpurun_out/final2/r02_grid3_256_ncu.txt 2>&1
rm -f gpurun_out/final2/r02_grid3_256.ncu-rep
for n in 128 256 384 512; do timeout 300 python scripts/bench_grid3.py $n 10 > gpurun_out/final2/grid3_$n.json 2> gpurun_out/final2/grid3_$n.err; echo "grid3 $n rc=$?"; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final2/r02_launches_8192.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/final2/launches.log 2>&1; echo "launch list rc=$?"
du -sh gpurun_out/final2
