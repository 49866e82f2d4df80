# round 2, call w: backward epilogues that absorb the channel-mask and dXc row kernels -- tests, A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py tests/test_gpu_round2.py tests/test_gpu_zfullsize.py tests/test_gpu_fused_attn.py -m gpu -q --timeout 300 --tb=line 2>&1 | grep -v "UserWarning\|^  " | cut -c1-700 | tail -30 > gpurun_out/r2w_tests.log
tail -12 gpurun_out/r2w_tests.log
for cfg in "default:" "noepi:DGSCT_NO_FUSED_BWD_EPI=1"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 200 python bench.py --no-eager --no-cpu-baseline --steps 15 > gpurun_out/r2w_bench_train_$name.json 2> gpurun_out/r2w_bench_train_$name.err
  env $envs timeout 300 python scripts/stage_table_bwd.py > gpurun_out/r2w_stage_bwd_$name.log 2>&1
  echo "== $name"; grep -E "^stage|bwd_channel|bwd_spatial|channel_query|spatial_query|TOTAL|weighted" gpurun_out/r2w_stage_bwd_$name.log
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2w_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        sm = d['stage_ms_per_step']
        print(f.split('r2w_bench_')[1][:-5], 'ms', round(d['ms_per_step'], 2), 'fwd', round((d.get('forward_only') or {}).get('ms_per_step', 0), 2), 'launches', d['gpu_launches'] // d['steps'], {k: round(v, 2) for k, v in sm.items() if k in ('bwd_channel', 'bwd_spatial', 'channel_query', 'spatial_query')})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-800:])
PY
