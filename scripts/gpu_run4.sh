mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q --timeout 120 2>&1 | tail -15 > gpurun_out/r4_gemm.log
tail -5 gpurun_out/r4_gemm.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -q --timeout 300 2>&1 | tail -30 > gpurun_out/r4_tests.log
tail -8 gpurun_out/r4_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r4_bench.log 2>&1
python - <<'PY'
import json
l=open('gpurun_out/r4_bench.log').read().strip().splitlines()[-1]
d=json.loads(l)
print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['achieved'])
print(d['stage_ms_per_step'])
PY
