#!/bin/bash
# round 2 (second session), 2 GPUs: loop-back tests of the multi-rank step (incl. pano_dist_step_host), the bench line with its parity check
set -u
mkdir -p gpurun_out/dist2
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/dist2/pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -3 gpurun_out/dist2/pytest_dist.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 5 --no-poisson > gpurun_out/dist2/bench_2gpu.json 2> gpurun_out/dist2/bench_2gpu.err
echo "bench 2gpu rc=$?"; tail -3 gpurun_out/dist2/bench_2gpu.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/dist2/bench_2gpu.json"))
    print("value", d["value"], "median ms", d["median_ms_per_step"])
    print("parity", d.get("parity_vs_one_gpu")); print("e2e", d.get("e2e"))
    print("speedup", d["value"] / d["one_gpu_same_workload"]["value"])
except Exception as e:
    print("no line", e)
PY
