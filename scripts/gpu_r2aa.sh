# round 2, call aa: spatial map beside gate_ln (auxiliary stream), dq split in one launch -- tests, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --tb=line 2>&1 | grep -v "UserWarning\|^  " | cut -c1-600 | tail -25 > gpurun_out/r2aa_tests.log
tail -8 gpurun_out/r2aa_tests.log
timeout 300 python bench.py --no-eager --no-cpu-baseline --steps 15 > gpurun_out/r2aa_bench_train.json 2> gpurun_out/r2aa_bench_train.err
timeout 300 python bench.py --mode fwd --no-eager --no-cpu-baseline > gpurun_out/r2aa_bench_fwd.json 2> gpurun_out/r2aa_bench_fwd.err
DGSCT_NO_SIDE=1 timeout 300 python bench.py --mode fwd --no-eager --no-cpu-baseline > gpurun_out/r2aa_bench_fwd_noside.json 2> gpurun_out/r2aa_bench_fwd_noside.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2aa_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('r2aa_bench_')[1][:-5], 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['ms_per_step'], 2), 'fwd', round((d.get('forward_only') or {}).get('ms_per_step', 0), 2), 'launches', d['gpu_launches'] // d['steps'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-800:])
PY
