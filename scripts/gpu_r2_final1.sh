#!/bin/bash
# final single-GPU records: full GPU test suite, the default bench line, the reference arm, ncu of the final advection kernel
set -u
mkdir -p gpurun_out/final
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final/smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final/bench_n1.json 2> gpurun_out/final/bench_n1.err; echo "bench rc=$?"
export PANO_BENCH_MIN_WARMUP=3
timeout 900 ncu --set full --clock-control none -f --import-source on -k regex:'k_advect' -s 4 -c 1 -o gpurun_out/final/r02_advect_4096 python bench.py --steps 2 --no-cpu --no-extra --grid 4096 --warmup 3 > gpurun_out/final/ncu_adv.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py gpurun_out/final/r02_advect_4096.ncu-rep > gpurun_out/final/r02_advect_4096_ncu.txt 2>&1
timeout 900 ncu --set full --clock-control none -f -k regex:'k_advect' -s 4 -c 1 -o gpurun_out/final/r02_advect_8192 python bench.py --steps 2 --no-cpu --no-extra --grid 8192 --warmup 3 > gpurun_out/final/ncu_adv8.log 2>&1
python scripts/ncu_summary.py gpurun_out/final/r02_advect_8192.ncu-rep > gpurun_out/final/r02_advect_8192_ncu.txt 2>&1
rm -f gpurun_out/final/r02_advect_8192.ncu-rep
unset PANO_BENCH_MIN_WARMUP
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final/bench_ref_n1.json 2> gpurun_out/final/bench_ref_n1.err; echo "reference rc=$?"
head -c 600 gpurun_out/final/bench_ref_n1.json; echo
