#!/bin/bash
# 2 GPUs: what the cross-GPU exchange adds to an 8192-wide, 1024-row slab per GPU (the per-GPU share of 8192^2 on 8 GPUs),
# and what the fence modes (option cg_fence) change.
set -u
mkdir -p gpurun_out
T0=$SECONDS
stamp() { echo "[t=$((SECONDS - T0))s] $*"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for F in 0 1 2 3; do
  timeout 120 $TR scripts/prof_slab.py 2048 8192 cg_dynamic=0 cg_fence=$F 2>&1 | grep "^rank"
done
timeout 120 $TR scripts/prof_slab.py 2048 8192 cg_dynamic=1 cg_fence=0 2>&1 | grep "^rank"
timeout 120 $TR scripts/prof_slab.py 2048 8192 cg_dynamic=1 cg_fence=3 2>&1 | grep "^rank"
stamp "2-GPU done"
for B in 0 1 2 4; do
  timeout 120 python scripts/prof_slab.py 1024 8192 cg_dynamic=1 cg_batch=$B 2>&1 | tail -1
done
stamp done
