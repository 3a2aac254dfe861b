mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "non_power_of_two or standalone_compute_stats or conv_path" 2>&1 | tail -25
