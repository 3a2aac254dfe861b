timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "3d or rfftn" 2>&1 | tail -2
for i in 1 2 3; do python tools/bench_configs.py --configs E --steps 64 | cut -c1-175; done
timeout 300 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum -k regex:"plane_fwd|lead_kernel|plane_inv|pass_d" -s 12 -c 3 --csv --log-file gpurun_out/t64_warm_v7.csv python tools/bench_configs.py --configs E --steps 8 > /dev/null 2>&1; grep gpu__time gpurun_out/t64_warm_v7.csv | cut -d, -f15-
