#!/bin/bash
set -u
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
: > gpurun_out/sanitizer_v7.txt
for TOOL in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $TOOL python scripts/sanitize_v7.py (B200, round 1, v7: push all-reduce, light fences, halo flags)" >> gpurun_out/sanitizer_v7.txt
  START=$SECONDS
  timeout 45 $SAN --tool $TOOL python scripts/sanitize_v7.py >> gpurun_out/sanitizer_v7.txt 2>&1; echo "-- $TOOL rc=$? in $((SECONDS - START)) s" | tee -a gpurun_out/sanitizer_v7.txt
done
grep -E "^==|SUMMARY|rc=" gpurun_out/sanitizer_v7.txt
