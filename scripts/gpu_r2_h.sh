#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2h_pytest.log
for g in 4096 8192; do for om in 1 0; do
  PANO_BENCH_MIN_WARMUP=5 timeout 600 python bench.py --grid $g --steps 20 --warmup 5 --no-cpu --no-extra --opt cg_order_mid=$om 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$g order_mid=$om', d['value'], d['median_ms_per_step'], d['roofline']['phase_ms'])"
done; done
