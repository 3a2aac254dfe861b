mkdir -p gpurun_out
( for n in 2 8 32 96; do echo "worlds $n"; LNX_T64_WINDOW=256 timeout 300 python tools/ab_config_e.py --reps 12 --worlds $n; done ) > gpurun_out/r2_scan_ab_small.jsonl 2>gpurun_out/r2_scan_ab_small.err
cut -c1-120 gpurun_out/r2_scan_ab_small.jsonl; tail -2 gpurun_out/r2_scan_ab_small.err
