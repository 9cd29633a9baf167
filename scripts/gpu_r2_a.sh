#!/bin/bash
# round 2, call A: GPU parity tests + the reworked bench line + launch list
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc >> gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/r2a_bench_n1.err; head -c 1500 gpurun_out/r2a_bench_n1.json
