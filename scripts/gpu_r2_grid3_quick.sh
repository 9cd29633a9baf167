set -u
mkdir -p gpurun_out/g3
timeout 600 python -m pytest tests/test_gpu_grid3.py -x -q -m gpu > gpurun_out/g3/pytest_b.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/g3/pytest_b.log
for opt in "" cg3_zc=8 cg3_zc=16 cg3_zc=32 cg3_zc=64 "cg_blocks_per_sm=1"; do timeout 300 python scripts/bench_grid3.py 256 5 $opt > gpurun_out/g3/bench_256_opt.json 2>&1; echo "opt=$opt"; python -c "import json;d=json.load(open('gpurun_out/g3/bench_256_opt.json'));print(d['phases']['cg'])"; done
timeout 300 python scripts/bench_grid3.py 128 5 > gpurun_out/g3/bench_128_opt.json 2>&1; python -c "import json;d=json.load(open('gpurun_out/g3/bench_128_opt.json'));print(d['phases']['cg'], d['cg_info'])"
timeout 600 ncu --set full --clock-control none -f -k regex:'k3_cg' -s 8 -c 1 -o gpurun_out/g3/r02_grid3_cg_256 python scripts/bench_grid3.py 256 2 > gpurun_out/g3/ncu.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py gpurun_out/g3/r02_grid3_cg_256.ncu-rep > gpurun_out/g3/r02_grid3_cg_256_ncu.txt 2>&1
rm -f gpurun_out/g3/r02_grid3_cg_256.ncu-rep
