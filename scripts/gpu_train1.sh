mkdir -p gpurun_out
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/rt_bench.log 2>&1
python - <<'PY'
import json
txt=open('gpurun_out/rt_bench.log').read()
l=txt.strip().splitlines()[-1]
try:
    d=json.loads(l)
    print('train ms', d['ms_per_step'], 'clips/s', d['value'], 'fwd', d.get('forward_only'), 'e2e', d['e2e'], 'launches', d['gpu_launches'])
    print('cpu', d['cpu_baseline'])
    print('roof', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['stack'])
except Exception as e:
    print(txt[-3000:])
PY
