mkdir -p gpurun_out
LNX_T64_PERSIST=1 timeout 300 python -m pytest tests -m gpu -x -q -k "3d or config_E or full_size_3d" 2>&1 | tail -3
( timeout 300 python tools/ab_config_e.py --reps 16
  LNX_T64_PERSIST=1 timeout 300 python tools/ab_config_e.py --reps 16 ) > gpurun_out/r2_persist_ab_v1.jsonl 2>gpurun_out/r2_persist_ab_v1.err
cut -c1-150 gpurun_out/r2_persist_ab_v1.jsonl; tail -2 gpurun_out/r2_persist_ab_v1.err
