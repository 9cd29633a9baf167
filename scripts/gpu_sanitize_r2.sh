#!/bin/bash
set -u
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
: > gpurun_out/sanitizer_r2.txt
for TOOL in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $TOOL python scripts/sanitize_r2.py (B200, round 2: k_advect_tma, k_cg_sr, k_cg_resident_sr, fused-halo loop-back step)" >> gpurun_out/sanitizer_r2.txt
  START=$SECONDS
  timeout 240 $SAN --tool $TOOL --print-limit 20 python scripts/sanitize_r2.py >> gpurun_out/sanitizer_r2.txt 2>&1; echo "-- $TOOL rc=$? in $((SECONDS - START)) s" | tee -a gpurun_out/sanitizer_r2.txt
done
grep -E "^==|SUMMARY|rc=|Race|hazard" gpurun_out/sanitizer_r2.txt | head -60
