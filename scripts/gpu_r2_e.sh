#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_pcg.py tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2e_pytest.log
for o in "cg_single_reduction=1" "cg_single_reduction=1 cg_order_mid=0"; do
  timeout 300 python scripts/prof_slab.py 1024 8192 $o 2>&1 | grep -E "^rank|rror" | tee -a gpurun_out/r2e_prof_slab.txt
  timeout 300 python scripts/prof_slab.py 4096 4096 $o 2>&1 | grep -E "^rank|rror" | tee -a gpurun_out/r2e_prof_slab.txt
done
for g in 4096 8192; do for sr in 1 0; do
  PANO_BENCH_MIN_WARMUP=5 timeout 600 python bench.py --grid $g --steps 10 --warmup 5 --no-cpu --no-extra --opt cg_single_reduction=$sr 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$g sr=$sr', d['value'], d['median_ms_per_step'], d['roofline']['phase_ms']['cg'])"
done; done
