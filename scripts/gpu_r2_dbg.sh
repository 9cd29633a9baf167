#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_fused.py -m gpu -x -q -k "advect_tma_bit_exact and 128-256-30.0-1" > gpurun_out/r2dbg_memcheck.log 2>&1
echo "rc=$?"; grep -E "Invalid|at 0x|by thread|Address|in .*\.cu|=========     at" gpurun_out/r2dbg_memcheck.log | head -40; tail -5 gpurun_out/r2dbg_memcheck.log
