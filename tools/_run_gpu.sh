# scratch driver of a gpurun call: full validation of the build in the tree (GPU tests, smoke, default bench line)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_gpu_tests_v12.log
cat gpurun_out/r2_gpu_tests_v12.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2_smoke_v12.log
python bench.py > gpurun_out/r2_bench_v14.json 2> gpurun_out/r2_bench_v14.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_v14.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])
for k, v in d.get('secondary', {}).items():
    print(k, v.get('value'), v.get('ms_per_step'), v.get('roofline', {}).get('frac'), v.get('error'))
PY
