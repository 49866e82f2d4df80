mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_parity.py -q --timeout 300 2>&1 | tail -12 > gpurun_out/r7_tests.log
tail -4 gpurun_out/r7_tests.log | cut -c1-300
timeout 300 python scripts/gemm_bench.py 2>&1 | grep -E "^conv"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r7_bench.log 2>&1
python - <<'PY'
import json
l=open('gpurun_out/r7_bench.log').read().strip().splitlines()[-1]
d=json.loads(l)
print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['achieved'])
print(d['stage_ms_per_step'])
PY
