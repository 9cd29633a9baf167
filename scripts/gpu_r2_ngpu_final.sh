#!/bin/bash
# final multi-GPU record of round 2: the bench line on N GPUs (parity check against the one-GPU run, one-GPU leg, 16384^2 Poisson leg)
set -u
N=${1:-8}
mkdir -p gpurun_out/finalN
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/finalN/bench_${N}gpu.json 2> gpurun_out/finalN/bench_${N}gpu.err
echo "bench ${N}gpu rc=$?"; tail -2 gpurun_out/finalN/bench_${N}gpu.err | cut -c1-200
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/finalN/bench_${N}gpu.json"))
    print("value", d["value"], "median ms", d["median_ms_per_step"], "phases", d["roofline"]["phase_ms"])
    print("one gpu", d.get("one_gpu_same_workload")); print("parity", d.get("parity_vs_one_gpu"))
    p = d.get("poisson_16384"); print("poisson", {k: p[k] for k in ("median_ms_per_solve", "value", "speedup_vs_one_gpu")} if p else None); print("e2e", d.get("e2e"))
    print("speedup", d["value"] / d["one_gpu_same_workload"]["value"], "clocks", d.get("clocks"))
except Exception as e:
    print("no line", e)
PY
