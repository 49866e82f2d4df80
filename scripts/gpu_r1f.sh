# call F: vectorised backward LN kernels, revised short-K recipe, capture diagnostics
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -40 > gpurun_out/f_tests.log
tail -3 gpurun_out/f_tests.log
for what in fwd fwdbwd fwdbwd_fused full; do
  timeout 200 python scripts/capture_debug.py global $what 2>&1 | grep -E "^\[|Error|error" | tail -4 >> gpurun_out/f_capture.log
done
timeout 200 python scripts/capture_debug.py thread_local full 2>&1 | grep -E "^\[|Error|error" | tail -4 >> gpurun_out/f_capture.log
timeout 200 python scripts/capture_debug.py relaxed full 2>&1 | grep -E "^\[|Error|error" | tail -4 >> gpurun_out/f_capture.log
cat gpurun_out/f_capture.log
timeout 300 python scripts/gemm_skinny.py "" > gpurun_out/f_skinny.log 2>&1
timeout 600 python bench.py --mode fwd > gpurun_out/f_bench_fwd.json 2> gpurun_out/f_bench_fwd.err
timeout 900 python bench.py > gpurun_out/f_bench_train.json 2> gpurun_out/f_bench_train.err
timeout 300 python scripts/stage_table_bwd.py > gpurun_out/f_stage_bwd.log 2>&1
tail -c 600 gpurun_out/f_bench_train.err
