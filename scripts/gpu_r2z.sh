# round 2, call z: fewer launches on the critical path (bn1 finalize + ReLU merged, BN-backward coefficients inline, ReLU mask +
# bn1 sums in the up-projection backward epilogue, fc row sums at pack time) -- tests, bench, stage table
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --tb=line 2>&1 | grep -v "UserWarning\|^  " | cut -c1-600 | tail -25 > gpurun_out/r2z_tests.log
tail -10 gpurun_out/r2z_tests.log
timeout 300 python bench.py --no-eager --no-cpu-baseline --steps 15 > gpurun_out/r2z_bench_train.json 2> gpurun_out/r2z_bench_train.err
DGSCT_NO_FUSED_BWD_EPI=1 timeout 300 python bench.py --no-eager --no-cpu-baseline --steps 15 > gpurun_out/r2z_bench_train_noepi.json 2> gpurun_out/r2z_bench_train_noepi.err
timeout 300 python bench.py --mode fwd --no-eager --no-cpu-baseline > gpurun_out/r2z_bench_fwd.json 2> gpurun_out/r2z_bench_fwd.err
timeout 300 python scripts/stage_table_bwd.py > gpurun_out/r2z_stage_bwd.log 2>&1
grep -E "^stage|bn_finalize|bwd_updown|misc|resample_fc|TOTAL|weighted" gpurun_out/r2z_stage_bwd.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2z_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        sm = d['stage_ms_per_step']
        print(f.split('r2z_bench_')[1][:-5], 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['ms_per_step'], 2), 'fwd', round((d.get('forward_only') or {}).get('ms_per_step', 0), 2), 'launches', d['gpu_launches'] // d['steps'], {k: round(v, 2) for k, v in sm.items() if k in ('bn_finalize', 'bwd_updown', 'misc', 'resample_fc')})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-800:])
PY
