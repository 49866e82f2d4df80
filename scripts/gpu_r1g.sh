# call G: split-K skinny backward products, graph-captured train step (bucketer fix), staged-store toggle on the fwd bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -30 > gpurun_out/g_tests.log
tail -3 gpurun_out/g_tests.log
timeout 200 python scripts/capture_debug.py global full 2>&1 | grep -E "^\[|Error|error" | tail -4 > gpurun_out/g_capture.log
cat gpurun_out/g_capture.log
timeout 600 python bench.py --mode fwd > gpurun_out/g_bench_fwd.json 2> gpurun_out/g_bench_fwd.err
DGSCT_NO_STAGE_OUT=1 timeout 600 python bench.py --mode fwd > gpurun_out/g_bench_fwd_nostage.json 2> gpurun_out/g_bench_fwd_nostage.err
timeout 900 python bench.py > gpurun_out/g_bench_train.json 2> gpurun_out/g_bench_train.err
timeout 300 python scripts/stage_table_bwd.py > gpurun_out/g_stage_bwd.log 2>&1
tail -c 600 gpurun_out/g_bench_train.err
