#!/bin/bash
# One gpurun call (round 1, v7): the full GPU suite, smoke, the default bench line, the ncu launch list and the
# 4096^2 / 128^2 lines.
# Everything is bounded by `timeout`; logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
T0=$SECONDS
stamp() { echo "[t=$((SECONDS - T0))s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvsmi.txt 2>&1
timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1; stamp "build rc=$?"
stamp "full suite"
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; stamp "tests rc=$?"; tail -6 gpurun_out/tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; stamp "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 400 python bench.py --steps 20 --warmup 20 > gpurun_out/bench_1024.json 2> gpurun_out/bench_1024.err; stamp "bench default rc=$?"; cat gpurun_out/bench_1024.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_1024.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-mg > gpurun_out/ncu_launch.log 2>&1; stamp "ncu rc=$?"
timeout 300 python bench.py --steps 5 --warmup 5 --n 4096 --no-cpu > gpurun_out/bench_4096.json 2> gpurun_out/bench_4096.err; stamp "bench 4096 rc=$?"; cat gpurun_out/bench_4096.json
timeout 200 python bench.py --steps 50 --warmup 50 --n 128 --no-cpu > gpurun_out/bench_128.json 2> gpurun_out/bench_128.err; stamp "bench 128 rc=$?"; cat gpurun_out/bench_128.json
stamp done
