#!/bin/bash
# 2 GPUs: halo-flag protocol of the cross-GPU reductions (option cg_xflags) -- loop-back parity, then A/B on the
# 1024 x 8192 slab per GPU, then the 8192^2 bench line.
set -u
mkdir -p gpurun_out
T0=$SECONDS
stamp() { echo "[t=$((SECONDS - T0))s] $*"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout 200 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/tests_dist_xflags1.log 2>&1; stamp "loop-back tests (xflags=1) rc=$?"; tail -2 gpurun_out/tests_dist_xflags1.log
PANO_OPT_cg_xflags=0 timeout 200 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/tests_dist_xflags0.log 2>&1; stamp "loop-back tests (xflags=0) rc=$?"; tail -2 gpurun_out/tests_dist_xflags0.log
for O in "cg_xflags=0" "cg_xflags=1" "cg_xflags=1 cg_fence=1" "cg_xflags=1 cg_fence=3"; do
  timeout 120 $TR scripts/prof_slab.py 2048 8192 cg_dynamic=0 $O 2>&1 | grep "^rank"
done
timeout 120 $TR scripts/prof_slab.py 2048 8192 cg_dynamic=1 cg_xflags=1 2>&1 | grep "^rank"
stamp "profiles done"
timeout 200 $TR bench.py --gpus 2 --steps 5 --warmup 5 --no-single > gpurun_out/bench_8192_2gpu.json 2> gpurun_out/bench_8192_2gpu.err; stamp "bench 2 GPUs rc=$?"; cat gpurun_out/bench_8192_2gpu.json | cut -c1-600
stamp done
