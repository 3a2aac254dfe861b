mkdir -p gpurun_out
python tools/ab_config_d.py --reps 12 --burst 3 --unrolls 8,16,32,64 2>&1 | grep '"pdl": 0' | tee gpurun_out/r2_t2k_graph_ab_v5.jsonl
