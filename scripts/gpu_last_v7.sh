#!/bin/bash
set -u
mkdir -p gpurun_out
T0=$SECONDS
timeout 60 python -m pytest tests/test_gpu_dist.py tests/test_gpu_pcg.py tests/test_gpu_step.py -m gpu -x -q > gpurun_out/tests_last.log 2>&1; echo "[t=$((SECONDS - T0))s] tests rc=$?"; tail -2 gpurun_out/tests_last.log
echo "== compute-sanitizer --tool racecheck python scripts/sanitize_v7.py (after slot_word: ring-slot words read by lane 0 and broadcast)" > gpurun_out/sanitizer_v7b.txt
timeout 40 /usr/local/cuda/bin/compute-sanitizer --tool racecheck python scripts/sanitize_v7.py >> gpurun_out/sanitizer_v7b.txt 2>&1; echo "[t=$((SECONDS - T0))s] racecheck rc=$?"; grep -E "SUMMARY" gpurun_out/sanitizer_v7b.txt
timeout 60 python bench.py --steps 5 --warmup 5 --n 4096 --no-cpu --no-mg > gpurun_out/bench_4096_last.json 2> gpurun_out/bench_4096_last.err; echo "[t=$((SECONDS - T0))s] bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_4096_last.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
