#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 60 scripts/probe/xgpu_probe 2 2000 2>&1 | tee gpurun_out/xgpu_probe_2.txt
for F in 0 4 5; do
  timeout 120 $TR scripts/prof_slab.py 2048 8192 cg_dynamic=0 cg_fence=$F 2>&1 | grep "^rank"
done
