# call D: tests of the new row kernels, skinny-GEMM ablation, bench fwd/train with graph-captured train step, 2-CTA/SM variant
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -6 > gpurun_out/d_tests.log
cat gpurun_out/d_tests.log
timeout 300 python scripts/gemm_skinny.py > gpurun_out/d_skinny.log 2>&1
timeout 600 python bench.py --mode fwd > gpurun_out/d_bench_fwd.json 2> gpurun_out/d_bench_fwd.err
DGSCT_TC_CTAS=2 timeout 600 python bench.py --mode fwd > gpurun_out/d_bench_fwd_cta2.json 2> gpurun_out/d_bench_fwd_cta2.err
timeout 600 python bench.py > gpurun_out/d_bench_train.json 2> gpurun_out/d_bench_train.err
tail -c 600 gpurun_out/d_bench_train.err
