mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_gpu_tests_v9.log
cat gpurun_out/r2_gpu_tests_v9.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2_smoke_v9.log
python bench.py > gpurun_out/r2_bench_v11.json 2> gpurun_out/r2_bench_v11.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_v11.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])
for k, v in d.get('secondary', {}).items():
    print(k, v.get('value'), v.get('ms_per_step'), v.get('roofline', {}).get('frac'), v.get('roofline', {}).get('hbm_convention', {}).get('engine_frac'), v.get('error'))
print(d['cpu_baseline'])
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_v11_reference.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench_v11_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_v11.csv python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > /dev/null 2>&1
grep -v "^==" gpurun_out/r2_launches_bench_v11.csv | cut -d, -f5,13- | sort | uniq -c | sort -rn | head -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lnx_world128_tm -s 2 -c 1 -f -o gpurun_out/r2_tm_v7 python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline --worlds 1184 --sim-steps 128 > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:lnx_world128_tm -s 2 -c 1 --csv --log-file gpurun_out/r2_tm_dram_traffic_v11.csv python bench.py --steps 1 --warmup 2 --no-secondary --no-cpu-baseline > /dev/null 2>&1
grep -v "^==" gpurun_out/r2_tm_dram_traffic_v11.csv | cut -d, -f5,13- | tail -3
