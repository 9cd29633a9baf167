#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a short bench, and the ncu launch list.
# Everything is bounded by `timeout`; logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvsmi.txt 2>&1
echo "== build" ; timeout 300 python __graft_entry__.py > gpurun_out/build.log 2>&1; echo "build rc=$?"
echo "== tests" ; timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > gpurun_out/tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/tests.log
echo "== smoke" ; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench 1024" ; timeout 600 python bench.py --steps 20 --warmup 20 > gpurun_out/bench_1024.json 2> gpurun_out/bench_1024.err; echo "bench rc=$?"; cat gpurun_out/bench_1024.json; tail -3 gpurun_out/bench_1024.err
echo "== bench 4096" ; timeout 600 python bench.py --steps 5 --warmup 5 --n 4096 --no-cpu > gpurun_out/bench_4096.json 2> gpurun_out/bench_4096.err; echo "bench rc=$?"; cat gpurun_out/bench_4096.json; tail -3 gpurun_out/bench_4096.err
echo "== bench 128" ; timeout 600 python bench.py --steps 50 --warmup 50 --n 128 --no-cpu > gpurun_out/bench_128.json 2> gpurun_out/bench_128.err; echo "bench rc=$?"; cat gpurun_out/bench_128.json
echo "== ncu launch list (1024^2, 2 steps after 3 warm-up)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_1024.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-mg > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?"
