#!/bin/bash
# round 2, Grid3d: parity tests, timing at 128^3 / 256^3, ncu of the 256^3 kernels
set -u
mkdir -p gpurun_out/g3
timeout 900 python -m pytest tests/test_gpu_grid3.py -x -q -m gpu > gpurun_out/g3/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/g3/pytest.log
for n in 128 256; do timeout 300 python scripts/bench_grid3.py $n 10 > gpurun_out/g3/bench_$n.json 2> gpurun_out/g3/bench_$n.err; echo "bench $n rc=$?"; python -c "import json;d=json.load(open('gpurun_out/g3/bench_$n.json'));print(d['ms_per_step'], d['cg_info'], {k:(round(v['ms'],4), round(v.get('frac_of_8000',0),3)) for k,v in d['phases'].items()})"; done
for opt in cg3_kernel=1 cg3_zc=4 cg3_zc=8 cg3_zc=12 cg3_zc=16 cg3_zc=32 "cg3_zc=16 cg_blocks_per_sm=1"; do timeout 300 python scripts/bench_grid3.py 256 5 $opt > gpurun_out/g3/bench_256_opt.json 2>&1; echo "$opt"; python -c "import json;d=json.load(open('gpurun_out/g3/bench_256_opt.json'));print(d['phases']['cg'])"; done
timeout 600 ncu --set full --clock-control none -f -k regex:'k3_' -s 40 -c 5 -o gpurun_out/g3/r02_grid3_256 python scripts/bench_grid3.py 256 2 > gpurun_out/g3/ncu.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py gpurun_out/g3/r02_grid3_256.ncu-rep > gpurun_out/g3/r02_grid3_256_ncu.txt 2>&1
rm -f gpurun_out/g3/r02_grid3_256.ncu-rep
