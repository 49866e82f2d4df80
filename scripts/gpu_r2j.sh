# round 2, call j: current per-shape stage table (forward+backward) and the combined tile-width / ring-depth A/B
mkdir -p gpurun_out
timeout 300 python scripts/stage_table_bwd.py > gpurun_out/r2j_stage_bwd.log 2>&1
timeout 300 python scripts/stage_table.py > gpurun_out/r2j_stage_fwd.log 2>&1
DGSCT_TC_BN_MAX=128 DGSCT_TC_DEEP_RING=0 timeout 200 python bench.py --no-eager --no-cpu-baseline --steps 15 > gpurun_out/r2j_bench_train_bn128_nodeep.json 2> gpurun_out/r2j_bench_train_bn128_nodeep.err
DGSCT_TC_BN_MAX=96 DGSCT_TC_DEEP_RING=0 timeout 200 python bench.py --no-eager --no-cpu-baseline --steps 15 > gpurun_out/r2j_bench_train_bn96_nodeep.json 2> gpurun_out/r2j_bench_train_bn96_nodeep.err
DGSCT_TC_BN_MAX=64 DGSCT_TC_DEEP_RING=0 timeout 200 python bench.py --no-eager --no-cpu-baseline --steps 15 > gpurun_out/r2j_bench_train_bn64_nodeep.json 2> gpurun_out/r2j_bench_train_bn64_nodeep.err
cat gpurun_out/r2j_stage_bwd.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2j_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('r2j_bench_')[1][:-5], 'ms', round(d['ms_per_step'], 2), 'fwd', round((d.get('forward_only') or {}).get('ms_per_step', 0), 2))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-500:])
PY
