# round 2, call i: validate the wide-row (C = 1536) LayerNorm-backward path and the deep-ring rule; A/B of tile width / ring depth
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_round2.py tests/test_gpu_zfullsize.py tests/test_gpu_gemm_tc.py tests/test_gpu_parity.py -m gpu -q --timeout 200 --tb=short 2>&1 | tail -30 > gpurun_out/r2i_tests.log
tail -4 gpurun_out/r2i_tests.log
timeout 200 python bench.py --no-eager --no-cpu-baseline --steps 15 > gpurun_out/r2i_bench_train_default.json 2> gpurun_out/r2i_bench_train_default.err
DGSCT_TC_DEEP_RING=0 timeout 200 python bench.py --no-eager --no-cpu-baseline --steps 15 > gpurun_out/r2i_bench_train_nodeepring.json 2> gpurun_out/r2i_bench_train_nodeepring.err
for bn in 128 192; do
  DGSCT_TC_BN_MAX=$bn timeout 200 python bench.py --no-eager --no-cpu-baseline --steps 15 > gpurun_out/r2i_bench_train_bn$bn.json 2> gpurun_out/r2i_bench_train_bn$bn.err
done
timeout 200 python bench.py --mode fwd --no-eager --no-cpu-baseline > gpurun_out/r2i_bench_fwd_default.json 2> gpurun_out/r2i_bench_fwd_default.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2i_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        sm = d['stage_ms_per_step']
        print(f.split('r2i_bench_')[1][:-5], 'ms', round(d['ms_per_step'], 2), 'fwd', round((d.get('forward_only') or {}).get('ms_per_step', 0), 2), {k: round(v, 2) for k, v in list(sm.items())[:14]})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-500:])
PY
