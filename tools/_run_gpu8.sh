mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_v11_n8.json 2> gpurun_out/r2_bench_v11_n8.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_v11_n8.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['scaling'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for k, v in d.get('secondary', {}).items():
    print(k, v.get('value'), v.get('ms_per_step'), v.get('error'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/check_sharded_multi_gpu.py > gpurun_out/r2_sharded_check_n8_v2.txt 2>&1; tail -6 gpurun_out/r2_sharded_check_n8_v2.txt
