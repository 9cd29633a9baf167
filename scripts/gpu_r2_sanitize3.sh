set -u
mkdir -p gpurun_out/san3
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_r2_grid3.py > gpurun_out/san3/$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|grid3|wide" gpurun_out/san3/$tool.log | tail -12
done
