mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "3d or config_E or rfftn or full_size_3d" 2>&1 | tail -5 > gpurun_out/r2_t64h_tests_v3.log
cat gpurun_out/r2_t64h_tests_v3.log
