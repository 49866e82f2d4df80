# round 2, call c: validate kflat / pair epilogues / exact fc.bias; bench A/Bs; ncu --set full of one backward call (2 shapes)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -25 > gpurun_out/r2c_tests.log
tail -4 gpurun_out/r2c_tests.log
timeout 300 python bench.py --no-eager > gpurun_out/r2c_bench_train.json 2> gpurun_out/r2c_bench_train.err
DGSCT_NO_KFLAT=1 timeout 300 python bench.py --no-eager --no-cpu-baseline > gpurun_out/r2c_bench_train_nokflat.json 2> gpurun_out/r2c_bench_train_nokflat.err
DGSCT_DW_EPI_SKIP=1 timeout 300 python bench.py --no-eager --no-cpu-baseline > gpurun_out/r2c_bench_train_dwskip.json 2> gpurun_out/r2c_bench_train_dwskip.err
DGSCT_NO_FC_FIRST=1 timeout 300 python bench.py --no-eager --no-cpu-baseline > gpurun_out/r2c_bench_train_nofcfirst.json 2> gpurun_out/r2c_bench_train_nofcfirst.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2c_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        sm = d['stage_ms_per_step']
        print(f.split('r2c_bench_')[1][:-5], 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['ms_per_step'], 2), 'fwd', round((d.get('forward_only') or {}).get('ms_per_step', 0), 2),
              {k: round(v, 2) for k, v in list(sm.items())[:10]})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-600:])
PY
for shape in aud_s2 vis_s0; do
  timeout 900 ncu --set full --clock-control none --profile-from-start off -o gpurun_out/r2c_full_$shape python scripts/one_call.py $shape 80 train > gpurun_out/r2c_ncu_$shape.log 2>&1
  ncu -i gpurun_out/r2c_full_$shape.ncu-rep --page raw --csv > gpurun_out/r2c_full_${shape}_raw.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/r2c_full_${shape}_raw.csv > gpurun_out/r2c_full_${shape}_summary.txt 2>&1
  rm -f gpurun_out/r2c_full_$shape.ncu-rep
  tail -3 gpurun_out/r2c_ncu_$shape.log
done
