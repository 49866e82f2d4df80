# round 2, call l: fused latent-token attention, forward + backward -- parity, fused-vs-chain check, A/B, stage tables
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py tests/test_gpu_round2.py tests/test_gpu_zfullsize.py -m gpu -q --timeout 300 --tb=line 2>&1 | grep -v "^tests.*UserWarning\|^  " | cut -c1-400 | tail -40 > gpurun_out/r2l_tests.log
tail -25 gpurun_out/r2l_tests.log
timeout 300 python scripts/fused_attn_check.py > gpurun_out/r2l_fused_check.log 2>&1; cut -c1-330 gpurun_out/r2l_fused_check.log
timeout 200 python bench.py --no-eager --no-cpu-baseline --steps 15 > gpurun_out/r2l_bench_train.json 2> gpurun_out/r2l_bench_train.err
DGSCT_NO_FUSED_ATTN_BWD=1 timeout 200 python bench.py --no-eager --no-cpu-baseline --steps 15 > gpurun_out/r2l_bench_train_nofusedbwd.json 2> gpurun_out/r2l_bench_train_nofusedbwd.err
timeout 200 python bench.py --mode fwd --no-eager --no-cpu-baseline > gpurun_out/r2l_bench_fwd.json 2> gpurun_out/r2l_bench_fwd.err
timeout 300 python scripts/stage_table.py > gpurun_out/r2l_stage_fwd.log 2>&1
timeout 300 python scripts/stage_table_bwd.py > gpurun_out/r2l_stage_bwd.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2l_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        sm = d['stage_ms_per_step']
        print(f.split('r2l_bench_')[1][:-5], 'ms', round(d['ms_per_step'], 2), 'fwd', round((d.get('forward_only') or {}).get('ms_per_step', 0), 2), {k: round(v, 2) for k, v in list(sm.items())[:12]})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-800:])
PY
grep -E "^stage|tok_update|x_update|TOTAL|weighted" gpurun_out/r2l_stage_fwd.log
grep -E "^stage|bwd_token|col_softmax|row_softmax|tok_update|x_update|TOTAL|weighted" gpurun_out/r2l_stage_bwd.log
