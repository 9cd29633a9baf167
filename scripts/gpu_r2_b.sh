#!/bin/bash
# round 2, call B: the single-reduction CG kernel -- parity tests, then timing against the two-reduction kernel
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pcg.py tests/test_gpu_dist.py tests/test_gpu_step.py -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2b_pytest.log
for g in 4096 8192; do
  for sr in 0 1; do
    timeout 600 python bench.py --grid $g --steps 10 --warmup 3 --no-cpu --no-extra --opt cg_single_reduction=$sr > gpurun_out/r2b_bench_${g}_sr$sr.json 2> gpurun_out/r2b_bench_${g}_sr$sr.err
    echo "bench $g sr=$sr rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2b_bench_${g}_sr$sr.json"))
    print(d["value"], d["median_ms_per_step"], d["roofline"]["kernel_ms"], d["cg_info_last_step"])
except Exception as e:
    print("no line", e)
PY
  done
done
