set -u
mkdir -p gpurun_out/wide
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q -m gpu > gpurun_out/wide/pytest0.log 2>&1; echo "pytest(default) rc=$?"; tail -2 gpurun_out/wide/pytest0.log
for wdt in 8 16; do PANO_OPT_fused_wide=$wdt timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_step.py -x -q -m gpu > gpurun_out/wide/pytest$wdt.log 2>&1; echo "pytest(wide=$wdt) rc=$?"; tail -2 gpurun_out/wide/pytest$wdt.log; done
for n in 4096 8192; do for wdt in 0 8 16; do echo "n=$n wide=$wdt"; PANO_OPT_fused_wide=$wdt timeout 300 python scripts/bench_kernels.py $n 2>&1 | grep -v advect_all; done; done
