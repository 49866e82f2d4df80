mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_parity.py -q --timeout 300 2>&1 | tail -8 | cut -c1-400
timeout 600 python scripts/stage_table_bwd.py 2>&1 | tail -34
