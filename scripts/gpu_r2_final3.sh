#!/bin/bash
# last single-GPU check of round 2: the whole GPU suite (with the 256^3 Grid3d test) and smoke
set -u
mkdir -p gpurun_out/final3
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final3/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final3/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final3/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final3/smoke.log | cut -c1-220
