# call H: weight-gradient products on a forked side stream (A/B against DGSCT_NO_SIDE)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -30 > gpurun_out/h_tests.log
tail -3 gpurun_out/h_tests.log
timeout 900 python bench.py > gpurun_out/h_bench_train.json 2> gpurun_out/h_bench_train.err
DGSCT_NO_SIDE=1 timeout 900 python bench.py > gpurun_out/h_bench_train_noside.json 2> gpurun_out/h_bench_train_noside.err
timeout 600 python bench.py --mode fwd > gpurun_out/h_bench_fwd.json 2> gpurun_out/h_bench_fwd.err
grep -E "^\[bench\]" gpurun_out/h_bench_train.err
