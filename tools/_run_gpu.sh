mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "2048_line2k" 2>&1 | tail -15 | tee gpurun_out/r2_graph_variant_test.log
