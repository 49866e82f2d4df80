# round 2, call g: final-state validation -- full GPU suite, smoke, bench lines of every workload, launch lists, targeted ncu --set full
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --tb=short 2>&1 | tail -60 > gpurun_out/r2g_tests.log
tail -6 gpurun_out/r2g_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; tail -3 gpurun_out/r2g_smoke.log
timeout 300 python bench.py > gpurun_out/r2g_bench_train.json 2> gpurun_out/r2g_bench_train.err
DGSCT_VEC_SIMT=1 timeout 300 python bench.py --no-eager --no-cpu-baseline > gpurun_out/r2g_bench_train_vecsimt.json 2> gpurun_out/r2g_bench_train_vecsimt.err
timeout 300 python bench.py --mode fwd --no-eager > gpurun_out/r2g_bench_fwd.json 2> gpurun_out/r2g_bench_fwd.err
for w in s4 avqa_t60 b224; do
  timeout 400 python bench.py --workload $w --no-cpu-baseline --steps 10 > gpurun_out/r2g_bench_train_$w.json 2> gpurun_out/r2g_bench_train_$w.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2g_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        sm = d['stage_ms_per_step']
        print(f.split('r2g_bench_')[1][:-5], 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['ms_per_step'], 2), 'fwd', round((d.get('forward_only') or {}).get('ms_per_step', 0), 2),
              'launches', d['gpu_launches'] // d['steps'], 'eager', (d.get('gpu_eager_baseline') or {}).get('ms_per_step'), 'roof', d['roofline']['stage'], round(d['roofline']['frac'], 3),
              {k: round(v, 2) for k, v in list(sm.items())[:8]})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-800:])
PY
DGSCT_BENCH_QUICK=1 DGSCT_BENCH_PROFILE=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2g_launches_train.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-eager > gpurun_out/r2g_ncu_train.log 2>&1
DGSCT_BENCH_QUICK=1 DGSCT_BENCH_PROFILE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2g_launches_fwd.csv python bench.py --mode fwd --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-eager > gpurun_out/r2g_ncu_fwd.log 2>&1
# targeted --set full: the weight-gradient GEMMs, the LayerNorm-backward row kernels and the bias-sum kernel of one stage-0 audio call
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:"EpiOut<0, 2|bwd_gate_ln_vec|bwd_out_vec|dv_bias_sums|EpiStore<0, 0, 0, 1>" -o gpurun_out/r2g_full_aud_s0 python scripts/one_call.py aud_s0 80 train > gpurun_out/r2g_ncu_full.log 2>&1
ncu -i gpurun_out/r2g_full_aud_s0.ncu-rep --page raw --csv > gpurun_out/r2g_full_aud_s0_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2g_full_aud_s0_raw.csv > gpurun_out/r2g_full_aud_s0_summary.txt 2>&1
rm -f gpurun_out/r2g_full_aud_s0.ncu-rep gpurun_out/r2g_full_aud_s0_raw.csv
tail -20 gpurun_out/r2g_full_aud_s0_summary.txt | cut -c1-200
