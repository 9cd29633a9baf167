#!/bin/bash
# 1 GPU: light fences (option cg_fence) in the streaming CG kernel -- parity incl. the loop-back multi-rank tests, then A/B.
set -u
mkdir -p gpurun_out
T0=$SECONDS
stamp() { echo "[t=$((SECONDS - T0))s] $*"; }
timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1; stamp "build rc=$?"
PANO_OPT_cg_fence=3 timeout 400 python -m pytest tests/test_gpu_dist.py tests/test_gpu_pcg.py tests/test_gpu_step.py -m gpu -x -q > gpurun_out/tests_fence3.log 2>&1; stamp "fence=3 tests rc=$?"; tail -3 gpurun_out/tests_fence3.log
timeout 300 python -m pytest tests/test_gpu_pcg.py -m gpu -x -q > gpurun_out/tests_pcg.log 2>&1; stamp "default pcg tests rc=$?"; tail -2 gpurun_out/tests_pcg.log
for F in 0 1 0 1; do
  timeout 120 python scripts/prof_slab.py 1024 8192 cg_dynamic=0 cg_fence=$F 2>&1 | tail -1
done
for F in 0 1; do
  timeout 120 python scripts/prof_slab.py 4096 4096 cg_fence=$F 2>&1 | tail -1
  timeout 120 python scripts/prof_slab.py 8192 8192 cg_fence=$F 2>&1 | tail -1
done
stamp done
