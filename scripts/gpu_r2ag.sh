# round 2, call ag: bench lines of the final tree (the driver's default invocation, forward)
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/r2ag_bench_train.json 2> gpurun_out/r2ag_bench_train.err
timeout 300 python bench.py --mode fwd --no-eager --no-cpu-baseline > gpurun_out/r2ag_bench_fwd.json 2> gpurun_out/r2ag_bench_fwd.err
python - <<'PY'
import json
for n in ('train', 'fwd'):
    d = json.loads(open(f'gpurun_out/r2ag_bench_{n}.json').read().strip().splitlines()[-1])
    print(n, 'ms', round(d['ms_per_step'], 2), 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'launches', d['gpu_launches'] // d['steps'], 'eager', (d.get('gpu_eager_baseline') or {}).get('ms_per_step'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'roof', d['roofline']['stage'], round(d['roofline']['frac'], 3), d['clocks'])
PY
