# first GPU call of round 2: validate the default-off experiments of round 1 and A/B them on the bench
# (DGSCT_EARLY_RELEASE, DGSCT_FUSE_BWD, DGSCT_FC_FIRST).  ~6 minutes of box time.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -4 > gpurun_out/r2a_tests.log
tail -2 gpurun_out/r2a_tests.log
timeout 300 python -m pytest tests -m gpu_experimental -q --timeout 120 2>&1 | tail -30 > gpurun_out/r2a_experimental.log
tail -4 gpurun_out/r2a_experimental.log
timeout 200 python bench.py > gpurun_out/r2a_bench_train_default.json 2> gpurun_out/r2a_bench_train_default.err
for flag in DGSCT_EARLY_RELEASE DGSCT_FUSE_BWD DGSCT_FC_FIRST; do
  env $flag=1 timeout 200 python bench.py > gpurun_out/r2a_bench_train_$flag.json 2> gpurun_out/r2a_bench_train_$flag.err
done
env DGSCT_EARLY_RELEASE=1 DGSCT_FUSE_BWD=1 DGSCT_FC_FIRST=1 timeout 200 python bench.py > gpurun_out/r2a_bench_train_all.json 2> gpurun_out/r2a_bench_train_all.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2a_bench_train_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('r2a_bench_train_')[1][:-5], 'train ms', round(d['ms_per_step'], 2), 'fwd ms', round(d['forward_only']['ms_per_step'], 2))
    except Exception as e:
        print(f, 'ERR', e)
PY
