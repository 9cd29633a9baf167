#!/bin/bash
set -u
mkdir -p gpurun_out
T0=$SECONDS
stamp() { echo "[t=$((SECONDS - T0))s] $*"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514"
timeout 200 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/tests_dist_v8.log 2>&1; stamp "loop-back tests rc=$?"; tail -2 gpurun_out/tests_dist_v8.log
for O in "cg_dynamic=0 cg_halo_first=1" "cg_dynamic=0 cg_halo_first=0" ""; do
  timeout 120 $TR scripts/prof_slab.py 2048 8192 $O 2>&1 | grep "^rank"
done
stamp done
