timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "3d or rfftn" 2>&1 | tail -2
for i in 1 2 3; do python tools/bench_configs.py --configs E --steps 64 | cut -c1-175; done
