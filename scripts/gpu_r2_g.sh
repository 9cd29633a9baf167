#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fused.py tests/test_gpu_dist.py -m gpu -x -q -k "advect or loopback_tma or step" > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2g_pytest.log
python scripts/time_advect.py 4096x4096 8192x8192 advect_border_first=1 | grep "kernel=0"
python scripts/time_advect.py 4096x4096 8192x8192 advect_border_first=0 | grep "kernel=0"
for g in 4096 8192; do
  PANO_BENCH_MIN_WARMUP=5 timeout 600 python bench.py --grid $g --steps 20 --warmup 5 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$g', d['value'], d['median_ms_per_step'], d['roofline']['phase_ms'])"
done
