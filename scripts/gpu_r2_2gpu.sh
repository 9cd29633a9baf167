#!/bin/bash
# round 2, 2 GPUs: the per-GPU slab of 8192^2 on 8 GPUs (2048 x 8192 over 2 ranks) under the protocol variants, then the bench line with its parity check
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for opts in "cg_single_reduction=0 dist_fused_halos=0" "cg_single_reduction=1" "cg_single_reduction=1 cg_halo_mid=0" "cg_single_reduction=1 cg_order_mid=0" "cg_single_reduction=1 cg_fence=1"; do
  timeout 300 $TR scripts/time_slab.py 2048 8192 $opts 2>&1 | grep -E "^rank|Error|error" | tee -a gpurun_out/r2_2gpu_slab.txt
done
timeout 300 python scripts/time_slab.py 1024 8192 cg_single_reduction=0 2>&1 | grep -E "^rank|rror" | tee -a gpurun_out/r2_2gpu_slab.txt
timeout 300 python scripts/time_slab.py 1024 8192 cg_single_reduction=1 2>&1 | grep -E "^rank|rror" | tee -a gpurun_out/r2_2gpu_slab.txt
PANO_BENCH_MIN_WARMUP=5 timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 5 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
echo "bench 2gpu rc=$?"; tail -3 gpurun_out/r2_bench_2gpu.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_2gpu.json"))
    print("value", d["value"], "median ms", d["median_ms_per_step"], "phases", d["roofline"]["phase_ms"])
    print("one gpu", d.get("one_gpu_same_workload")); print("parity", d.get("parity_vs_one_gpu")); print("poisson", d.get("poisson_16384")); print("e2e", d.get("e2e"))
except Exception as e:
    print("no line", e)
PY
