# round 2, call h: A/B of the GEMM grid cap (share of CTA slots one GEMM may take while several streams run side by side)
mkdir -p gpurun_out
for pct in 100 75 50 35; do
  DGSCT_TC_SLOT_PCT=$pct timeout 200 python bench.py --no-eager --no-cpu-baseline --steps 15 > gpurun_out/r2h_bench_train_pct$pct.json 2> gpurun_out/r2h_bench_train_pct$pct.err
done
DGSCT_TC_SLOT_PCT=50 timeout 200 python bench.py --mode fwd --no-eager --no-cpu-baseline > gpurun_out/r2h_bench_fwd_pct50.json 2> gpurun_out/r2h_bench_fwd_pct50.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2h_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('r2h_bench_')[1][:-5], 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['ms_per_step'], 2), 'fwd', round((d.get('forward_only') or {}).get('ms_per_step', 0), 2))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-500:])
PY
