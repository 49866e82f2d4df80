# round 2, call af: last check of the committed tree -- GPU suite, smoke, default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -3 | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --no-eager > gpurun_out/r2af_bench_train.json 2> gpurun_out/r2af_bench_train.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2af_bench_train.json').read().strip().splitlines()[-1])
print('train ms', round(d['ms_per_step'], 2), 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'launches', d['gpu_launches'] // d['steps'], 'cpu', d['cpu_baseline']['value'], d['clocks'])
PY
