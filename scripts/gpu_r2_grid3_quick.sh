set -u
mkdir -p gpurun_out/g3
timeout 900 python -m pytest tests/test_gpu_grid3.py -x -q -m gpu > gpurun_out/g3/pytest_g.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/g3/pytest_g.log
for n in 128 256 384 512; do timeout 300 python scripts/bench_grid3.py $n 10 > gpurun_out/g3/bench_$n.json 2> gpurun_out/g3/bench_$n.err; echo "bench $n rc=$?"; python -c "import json;d=json.load(open('gpurun_out/g3/bench_$n.json'));print(d['ms_per_step'], d['mcell_steps_per_s'], sum(d['cg_applies_per_timed_step']), {k:(round(v['ms'],4), round(v.get('frac_of_8000',0),3)) for k,v in d['phases'].items()})"; done
for opt in cg3_zc=8 cg3_zc=32; do timeout 300 python scripts/bench_grid3.py 256 10 $opt > gpurun_out/g3/bench_256_opt.json 2>&1; python -c "import json;d=json.load(open('gpurun_out/g3/bench_256_opt.json'));print('$opt', d['phases']['cg']['ms'], d['phases']['cg']['frac_of_8000'], sum(d['cg_applies_per_timed_step']))"; done
