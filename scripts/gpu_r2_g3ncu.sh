set -u
mkdir -p gpurun_out/final4
for n in 128 256 384 512; do timeout 300 python scripts/bench_grid3.py $n 10 > gpurun_out/final4/grid3_$n.json 2> gpurun_out/final4/grid3_$n.err; echo "grid3 $n rc=$?"; python -c "import json;d=json.load(open('gpurun_out/final4/grid3_$n.json'));print(d['ms_per_step'], d['mcell_steps_per_s'], d['cg_applies_per_timed_step'], {k:(round(v['ms'],4), round(v.get('frac_of_8000',0),3)) for k,v in d['phases'].items()})"; done
