mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_gpu_tests_v11.log
cat gpurun_out/r2_gpu_tests_v11.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2_smoke_v11.log
python bench.py > gpurun_out/r2_bench_v13.json 2> gpurun_out/r2_bench_v13.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_v13.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])
for k, v in d.get('secondary', {}).items():
    print(k, v.get('value'), v.get('ms_per_step'), v.get('roofline', {}).get('frac'), v.get('roofline', {}).get('hbm_convention', {}).get('engine_frac'), v.get('error'))
PY
