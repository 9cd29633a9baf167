#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dist.py tests/test_gpu_pcg.py -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2f_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for opts in "cg_single_reduction=1" "cg_single_reduction=1 cg_early_load=0" "cg_single_reduction=0 dist_fused_halos=0"; do
  timeout 300 $TR scripts/time_slab.py 2048 8192 $opts 2>&1 | grep -E "^rank|Error|error" | tee -a gpurun_out/r2f_2gpu_slab.txt
done
