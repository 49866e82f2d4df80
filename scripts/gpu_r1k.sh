# call K: lean producer / MMA-issuer loops -- tests, probe (1 CTA/SM and 2), bench lines
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q --timeout 100 2>&1 | tail -8 > gpurun_out/k_tests.log
tail -2 gpurun_out/k_tests.log
timeout 200 python bench.py > gpurun_out/k_bench_train.json 2> gpurun_out/k_bench_train.err
timeout 100 python bench.py --mode fwd > gpurun_out/k_bench_fwd.json 2> gpurun_out/k_bench_fwd.err
timeout 60 python scripts/gemm_pipe_probe.py > gpurun_out/k_pipe_probe.log 2>&1
