# round 2, call d: full GPU suite (K != 32 tensor-core path, f3 heads, per-CTA BN statistics, light row geometry, aux streams),
# bench A/Bs, launch list of one train step
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -40 > gpurun_out/r2d_tests.log
tail -6 gpurun_out/r2d_tests.log
timeout 300 python bench.py --no-eager > gpurun_out/r2d_bench_train.json 2> gpurun_out/r2d_bench_train.err
DGSCT_ROW_HEAVY=1 timeout 300 python bench.py --no-eager --no-cpu-baseline > gpurun_out/r2d_bench_train_rowheavy.json 2> gpurun_out/r2d_bench_train_rowheavy.err
DGSCT_NO_SIDE=1 timeout 300 python bench.py --no-eager --no-cpu-baseline > gpurun_out/r2d_bench_train_noside.json 2> gpurun_out/r2d_bench_train_noside.err
timeout 300 python bench.py --mode fwd --no-eager --no-cpu-baseline > gpurun_out/r2d_bench_fwd.json 2> gpurun_out/r2d_bench_fwd.err
timeout 400 python bench.py --workload avqa_t60 --no-cpu-baseline --no-eager --steps 10 > gpurun_out/r2d_bench_train_avqa_t60.json 2> gpurun_out/r2d_bench_train_avqa_t60.err
timeout 400 python bench.py --workload s4 --no-cpu-baseline --no-eager --steps 10 > gpurun_out/r2d_bench_train_s4.json 2> gpurun_out/r2d_bench_train_s4.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2d_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        sm = d['stage_ms_per_step']
        print(f.split('r2d_bench_')[1][:-5], 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['ms_per_step'], 2), 'fwd', round((d.get('forward_only') or {}).get('ms_per_step', 0), 2),
              'launches', d['gpu_launches'] // d['steps'], {k: round(v, 2) for k, v in list(sm.items())[:12]})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-800:])
PY
DGSCT_BENCH_QUICK=1 DGSCT_BENCH_PROFILE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2d_launches_train.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-eager > gpurun_out/r2d_ncu_train.log 2>&1
tail -2 gpurun_out/r2d_ncu_train.log
