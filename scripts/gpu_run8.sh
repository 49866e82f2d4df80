mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py -q --timeout 300 -s 2>&1 | tail -150 > gpurun_out/r8_bwd.log
grep -E "^(FAILED|vis|aud|E   |[0-9]+ (passed|failed))" gpurun_out/r8_bwd.log | cut -c1-1200 | tail -40
