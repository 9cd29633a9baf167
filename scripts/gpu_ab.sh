#!/bin/bash
# A/B runs of bench.py with library options; usage: gpu_ab.sh N "opt1=a opt2=b" "opt1=c" ...
set -u
mkdir -p gpurun_out
N=$1; shift
K=20; [ "$N" -ge 4096 ] && K=5; [ "$N" -le 256 ] && K=50
i=0
for OPTS in "$@"; do
  ARGS=""; for o in $OPTS; do ARGS="$ARGS --opt $o"; done
  timeout 600 python bench.py --steps $K --warmup $K --n $N --no-cpu $ARGS > gpurun_out/ab_${N}_$i.json 2> gpurun_out/ab_${N}_$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_${N}_$i.json")); r=d["roofline"]
    print("N=$N [$OPTS] value=%.1f ms/step=%.3f cg_ms=%.3f alg_GB/s=%.0f" % (d["value"], d["ms_per_step"], r["kernel_ms"], r["achieved"]))
except Exception as e:
    print("N=$N [$OPTS] failed", e); print(open("gpurun_out/ab_${N}_$i.err").read()[-1500:])
PY
  i=$((i+1))
done
