mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -s 2>&1 | tail -60 > gpurun_out/r3_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r3_bench.log 2>&1
tail -40 gpurun_out/r3_tests.log; tail -c 2500 gpurun_out/r3_bench.log
