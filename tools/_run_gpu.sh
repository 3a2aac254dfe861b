mkdir -p gpurun_out
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r2_bench_v13_n$n.json 2> gpurun_out/r2_bench_v13_n$n.err
python - $n <<'PY'
import json, sys
d = json.loads(open('gpurun_out/r2_bench_v13_n%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
print(d['n_gpus'], d['scaling'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for k, v in d.get('secondary', {}).items():
    print(k, v.get('value'), v.get('ms_per_step'), v.get('error'))
PY
done
