mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/rb_bench.log 2>&1
python - <<'PY'
import json
l=open('gpurun_out/rb_bench.log').read().strip().splitlines()[-1]
try:
    d=json.loads(l)
    print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['achieved'], d['gpu_launches'])
    print(d['stage_ms_per_step'])
except Exception as e:
    print(open('gpurun_out/rb_bench.log').read()[-3000:])
PY
