# round 2, call r: lean per-frame-vector linear kernel -- tests, A/B against the tcgen05 GEMMs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py tests/test_gpu_round2.py tests/test_gpu_zfullsize.py tests/test_gpu_fused_attn.py -m gpu -q --timeout 300 --tb=line 2>&1 | grep -v "UserWarning\|^  " | cut -c1-600 | tail -30 > gpurun_out/r2r_tests.log
tail -12 gpurun_out/r2r_tests.log
for cfg in "default:" "nolean:DGSCT_VEC_LEAN=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 200 python bench.py --no-eager --no-cpu-baseline --steps 15 > gpurun_out/r2r_bench_train_$name.json 2> gpurun_out/r2r_bench_train_$name.err
done
timeout 300 python scripts/stage_table_bwd.py > gpurun_out/r2r_stage_bwd.log 2>&1
grep -E "^stage|small|TOTAL|weighted" gpurun_out/r2r_stage_bwd.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2r_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        sm = d['stage_ms_per_step']
        print(f.split('r2r_bench_')[1][:-5], 'ms', round(d['ms_per_step'], 2), 'fwd', round((d.get('forward_only') or {}).get('ms_per_step', 0), 2), {k: round(v, 2) for k, v in sm.items() if k in ('small_linear', 'bwd_small', 'bwd_token')})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-800:])
PY
