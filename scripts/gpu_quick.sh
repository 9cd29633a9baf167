#!/bin/bash
# tests + three bench sizes (no CPU leg); logs in gpurun_out/
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/tests.log
for N in ${SIZES:-1024 4096 128}; do
  K=20; [ "$N" -ge 4096 ] && K=5; [ "$N" -le 256 ] && K=50
  timeout 600 python bench.py --steps $K --warmup $K --n $N --no-cpu ${BENCH_ARGS:-} > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err; echo "bench $N rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$N.json"))
    r=d["roofline"]
    print("N=$N value=%.1f ms/step=%.3f cg_ms=%.3f cg_alg_GB/s=%.0f frac=%.3f e2e=%.1f iters=%d launches=%d" % (d["value"], d["ms_per_step"], r["kernel_ms"], r["achieved"], r["frac"], d["e2e"]["value"], d["config"]["cg_iterations_per_step"], d["gpu_launches"]))
    print("   phases", {k: round(v,4) for k,v in r["phase_ms"].items()}, d["clocks"])
except Exception as e:
    print("bench $N failed:", e); print(open("gpurun_out/bench_$N.err").read()[-2000:])
PY
done
