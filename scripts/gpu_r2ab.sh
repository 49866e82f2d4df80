# round 2, call ab: FINAL build -- full GPU suite, smoke, the driver's bench invocations, all workloads, launch lists, reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --tb=short 2>&1 | tail -30 | cut -c1-300 > gpurun_out/r2ab_tests.log
tail -4 gpurun_out/r2ab_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ab_smoke.log 2>&1; tail -5 gpurun_out/r2ab_smoke.log
timeout 400 python bench.py > gpurun_out/r2ab_bench_train.json 2> gpurun_out/r2ab_bench_train.err
timeout 300 python bench.py --mode fwd --no-eager --no-cpu-baseline > gpurun_out/r2ab_bench_fwd.json 2> gpurun_out/r2ab_bench_fwd.err
for w in s4 avqa_t60 b224; do
  timeout 400 python bench.py --workload $w --no-cpu-baseline --no-eager --steps 10 > gpurun_out/r2ab_bench_train_$w.json 2> gpurun_out/r2ab_bench_train_$w.err
done
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2ab_bench_ref.json 2> gpurun_out/r2ab_bench_ref.err
DGSCT_BENCH_QUICK=1 DGSCT_BENCH_PROFILE=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2ab_launches_train.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-eager > gpurun_out/r2ab_ncu_train.log 2>&1
DGSCT_BENCH_QUICK=1 DGSCT_BENCH_PROFILE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2ab_launches_fwd.csv python bench.py --mode fwd --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-eager > gpurun_out/r2ab_ncu_fwd.log 2>&1
wc -l gpurun_out/r2ab_launches_train.csv gpurun_out/r2ab_launches_fwd.csv
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2ab_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        if d.get('impl') == 'reference':
            print('reference', d.get('value'), d.get('unit'), d.get('ms_per_step'), d.get('cpu_baseline'))
            continue
        print(f.split('r2ab_bench_')[1][:-5], 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['ms_per_step'], 2), 'fwd', round((d.get('forward_only') or {}).get('ms_per_step', 0), 2),
              'launches', d['gpu_launches'] // d['steps'], 'eager', (d.get('gpu_eager_baseline') or {}).get('ms_per_step'), 'cpu', (d.get('cpu_baseline') or {}).get('value'),
              'roof', d['roofline']['stage'], round(d['roofline']['frac'], 3), d['clocks'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-800:])
PY
