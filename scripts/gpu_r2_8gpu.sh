#!/bin/bash
# round 2, 8 GPUs: 8192^2 strong scaling -- protocol variants (time_slab), then the bench line (parity check, one-GPU leg, 16384^2 Poisson leg)
set -u
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for opts in "cg_single_reduction=1" "cg_single_reduction=0 dist_fused_halos=0" "cg_single_reduction=1 cg_early_load=0 cg_halo_mid=0"; do
  timeout 300 $TR scripts/time_slab.py 8192 8192 $opts 2>&1 | grep -E "^rank|Error|error" | tee -a gpurun_out/r2_${N}gpu_slab.txt
done
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
echo "bench ${N}gpu rc=$?"; tail -3 gpurun_out/r2_bench_${N}gpu.err | cut -c1-300
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_${N}gpu.json"))
    print("value", d["value"], "median ms", d["median_ms_per_step"], "phases", d["roofline"]["phase_ms"])
    print("one gpu", d.get("one_gpu_same_workload")); print("parity", d.get("parity_vs_one_gpu")); print("poisson", d.get("poisson_16384")); print("e2e", d.get("e2e"))
    print("speedup", d["value"] / d["one_gpu_same_workload"]["value"])
except Exception as e:
    print("no line", e)
PY
