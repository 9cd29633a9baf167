set -u
mkdir -p gpurun_out/g3
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/g3/pytest_e.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/g3/pytest_e.log
for n in 128 256 512; do timeout 300 python scripts/bench_grid3.py $n 6 > gpurun_out/g3/bench_$n.json 2> gpurun_out/g3/bench_$n.err; echo "bench $n rc=$?"; python -c "import json;d=json.load(open('gpurun_out/g3/bench_$n.json'));print(d['ms_per_step'], d['mcell_steps_per_s'], d['cg_info'], {k:(round(v['ms'],4), round(v.get('frac_of_8000',0),3)) for k,v in d['phases'].items()})"; done
for zc in 8 16 32 64; do timeout 300 python scripts/bench_grid3.py 256 6 cg3_zc=$zc > gpurun_out/g3/bench_256_opt.json 2>&1; python -c "import json;d=json.load(open('gpurun_out/g3/bench_256_opt.json'));print('zc=$zc', d['phases']['cg']['ms'], d['phases']['cg']['frac_of_8000'])"; done
timeout 600 ncu --set full --clock-control none -f -k regex:'k3_cg' -s 8 -c 1 -o gpurun_out/g3/r02_grid3_cg_256 python scripts/bench_grid3.py 256 2 > gpurun_out/g3/ncu.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py gpurun_out/g3/r02_grid3_cg_256.ncu-rep > gpurun_out/g3/r02_grid3_cg_256_ncu.txt 2>&1
rm -f gpurun_out/g3/r02_grid3_cg_256.ncu-rep
