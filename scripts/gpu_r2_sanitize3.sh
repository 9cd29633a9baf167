set -u
mkdir -p gpurun_out/san3
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_r2_grid3.py > gpurun_out/san3/$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Barrier error" gpurun_out/san3/$tool.log | sort | uniq -c | tail -6
done
timeout 900 python -m pytest tests/test_gpu_grid3.py tests/test_gpu_host_cpp.py -x -q -m gpu 2>&1 | tail -2
for n in 128 256 384 512; do timeout 300 python scripts/bench_grid3.py $n 10 > gpurun_out/san3/grid3_$n.json 2> gpurun_out/san3/grid3_$n.err; python -c "import json;d=json.load(open('gpurun_out/san3/grid3_$n.json'));print($n, d['ms_per_step'], d['mcell_steps_per_s'], sum(d['cg_applies_per_timed_step']), {k:(round(v['ms'],4), round(v.get('frac_of_8000',0),3)) for k,v in d['phases'].items()})"; done
timeout 600 ncu --set full --clock-control none -f -k regex:'k3_advect|k3_neg_div|k3_cg|k3_project' -s 32 -c 4 -o gpurun_out/san3/r02_grid3_256 python scripts/bench_grid3.py 256 2 > gpurun_out/san3/ncu.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py gpurun_out/san3/r02_grid3_256.ncu-rep > gpurun_out/san3/r02_grid3_256_ncu.txt 2>&1
rm -f gpurun_out/san3/r02_grid3_256.ncu-rep
