# round 2, call b: full GPU test-suite after the kickoff changes, gradient error tables, bench lines of every workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -25 > gpurun_out/r2b_tests.log
tail -5 gpurun_out/r2b_tests.log
for v in ave s4 avqa; do timeout 300 python scripts/grad_table.py $v > gpurun_out/r2b_grad_table_$v.txt 2>&1; done
timeout 300 python bench.py > gpurun_out/r2b_bench_train.json 2> gpurun_out/r2b_bench_train.err
timeout 300 python bench.py --torch-adam --no-eager --no-cpu-baseline > gpurun_out/r2b_bench_train_torchadam.json 2> gpurun_out/r2b_bench_train_torchadam.err
timeout 300 python bench.py --mode fwd --no-eager > gpurun_out/r2b_bench_fwd.json 2> gpurun_out/r2b_bench_fwd.err
for w in s4 avqa_t60 b224; do
  timeout 400 python bench.py --workload $w --no-cpu-baseline --steps 10 > gpurun_out/r2b_bench_train_$w.json 2> gpurun_out/r2b_bench_train_$w.err
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2b_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('r2b_bench_')[1][:-5], 'ms', round(d['ms_per_step'], 2), 'e2e ms', d['e2e'].get('ms_per_step'), 'fwd', (d.get('forward_only') or {}).get('ms_per_step'),
              'roof', (d.get('roofline') or {}).get('stage'), round((d.get('roofline') or {}).get('frac', 0), 3), 'eager', (d.get('gpu_eager_baseline') or {}).get('ms_per_step'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-600:])
PY
