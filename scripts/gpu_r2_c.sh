#!/bin/bash
# round 2, call C: tests of the changed kernels, then full bench lines (with the 4096^2 / 1024^2 / 128^2 legs) for both CG arrangements
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fused.py tests/test_gpu_dist.py tests/test_gpu_step.py -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2c_pytest.log
for sr in 1 0; do
  PANO_BENCH_MIN_WARMUP=5 timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu --opt cg_single_reduction=$sr > gpurun_out/r2c_bench_sr$sr.json 2> gpurun_out/r2c_bench_sr$sr.err
  echo "bench sr=$sr rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2c_bench_sr$sr.json"))
    print(d["value"], d["median_ms_per_step"], d["roofline"]["phase_ms"])
    k=d["kernels_4096"]; print("4096:", k["value"], {n:(round(v["ms"],4), round(v.get("frac_of_8000",0),3)) for n,v in k["kernels"].items()})
    print("1024:", d["config_1024"]["value"], d["config_1024"]["median_ms_per_step"], d["config_1024"]["kernel"], "128:", d["config_128"]["value"], d["config_128"]["median_ms_per_step"])
    print("e2e", d["e2e"]["value"])
except Exception as e:
    print("no line", e)
PY
  tail -2 gpurun_out/r2c_bench_sr$sr.err
done
