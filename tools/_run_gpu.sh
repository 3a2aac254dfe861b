mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "2048 or large_2d or config_D" 2>&1 | tail -5 > gpurun_out/r2_t2k_tests_v2.log
cat gpurun_out/r2_t2k_tests_v2.log
( python tools/bench_configs.py --configs D --t2k-pairs; python tools/bench_configs.py --configs D; python tools/bench_configs.py --configs D --t2k-pairs; python tools/bench_configs.py --configs D ) 2>gpurun_out/r2_t2k_ab_v2.err | cut -c1-300 > gpurun_out/r2_t2k_ab_v2.txt
cat gpurun_out/r2_t2k_ab_v2.txt
ncu --set full --clock-control none --import-source on -k regex:"rows1_inv|lead_kernel" -s 40 -c 2 -f -o gpurun_out/r2_t2k_v2 python tools/bench_configs.py --configs D --steps 32 > gpurun_out/r2_t2k_ncu_v2.log 2>&1
tail -2 gpurun_out/r2_t2k_ncu_v2.log
