# call E: new GEMM recipe (resident B, staged stores, flattened frames, 2 CTAs/SM on short K), saved activations, fused grads
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q --timeout 120 2>&1 | tail -25 > gpurun_out/e_gemm.log
tail -3 gpurun_out/e_gemm.log
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -40 > gpurun_out/e_tests.log
tail -3 gpurun_out/e_tests.log
if grep -q "failed" gpurun_out/e_tests.log gpurun_out/e_gemm.log; then
  for v in DGSCT_NO_BRES DGSCT_NO_STAGE_OUT DGSCT_NO_FLATTEN; do
    echo "== $v" >> gpurun_out/e_bisect.log
    env $v=1 timeout 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_parity.py tests/test_gpu_backward.py -q --timeout 120 2>&1 | tail -12 >> gpurun_out/e_bisect.log
  done
  echo "== all off" >> gpurun_out/e_bisect.log
  env DGSCT_NO_BRES=1 DGSCT_NO_STAGE_OUT=1 DGSCT_NO_FLATTEN=1 timeout 600 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -12 >> gpurun_out/e_bisect.log
fi
timeout 300 python scripts/gemm_skinny.py "" > gpurun_out/e_skinny.log 2>&1
timeout 600 python bench.py --mode fwd > gpurun_out/e_bench_fwd.json 2> gpurun_out/e_bench_fwd.err
timeout 900 python bench.py > gpurun_out/e_bench_train.json 2> gpurun_out/e_bench_train.err
tail -c 800 gpurun_out/e_bench_train.err
