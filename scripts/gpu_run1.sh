mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/r1_gpu.txt 2>&1
lscpu | head -20 >> gpurun_out/r1_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -60 > gpurun_out/r1_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r1_smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r1_bench.log 2>&1
tail -5 gpurun_out/r1_tests.log; tail -3 gpurun_out/r1_smoke.log; tail -c 1500 gpurun_out/r1_bench.log
