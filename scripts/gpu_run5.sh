mkdir -p gpurun_out
timeout 300 python scripts/gemm_ablate.py 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_parity.py -q --timeout 300 2>&1 | tail -12 > gpurun_out/r5_tests.log
tail -6 gpurun_out/r5_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r5_bench.log 2>&1
python - <<'PY'
import json
l=open('gpurun_out/r5_bench.log').read().strip().splitlines()[-1]
d=json.loads(l)
print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['achieved'])
print(d['stage_ms_per_step'])
PY
