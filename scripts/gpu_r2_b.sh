#!/bin/bash
# round 2, call B: the single-reduction CG kernel and the TMA advection kernel -- parity tests, then timing against their predecessors
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fused.py tests/test_gpu_pcg.py tests/test_gpu_dist.py tests/test_gpu_step.py -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2b_pytest.log
for cfg in "4096 cg_single_reduction=0 advect_kernel=3" "4096 cg_single_reduction=1 advect_kernel=4" "4096 cg_single_reduction=1 advect_kernel=4 --opt advect_dynamic=0" "8192 cg_single_reduction=0 advect_kernel=3" "8192 cg_single_reduction=1 advect_kernel=4"; do
  set -- $cfg
  g=$1; o1=$2; o2=$3; shift 3
  tag="${g}_${o1#*=}_${o2#*=}$(echo $* | tr -d ' =-')"
  timeout 600 python bench.py --grid $g --steps 10 --warmup 3 --no-cpu --no-extra --opt $o1 --opt $o2 $* > gpurun_out/r2b_bench_$tag.json 2> gpurun_out/r2b_bench_$tag.err
  echo "bench $cfg rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2b_bench_$tag.json"))
    print(d["value"], d["median_ms_per_step"], d["roofline"]["phase_ms"], d["cg_info_last_step"])
except Exception as e:
    print("no line", e)
PY
  tail -2 gpurun_out/r2b_bench_$tag.err
done
