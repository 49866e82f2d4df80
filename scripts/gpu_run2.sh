mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q --timeout 120 2>&1 | tail -40 > gpurun_out/r2_gemm_tc.log
tail -25 gpurun_out/r2_gemm_tc.log
