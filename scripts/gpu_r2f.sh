# round 2, call f (2 GPUs): N=2 bench after the profiling-deadlock fix (short timeouts), failing tests re-run, N=1 bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_backward.py tests/test_gpu_parity.py -m gpu -q --timeout 200 --tb=short 2>&1 | tail -40 > gpurun_out/r2f_tests.log
tail -6 gpurun_out/r2f_tests.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --no-eager > gpurun_out/r2f_bench_train_n2.json 2> gpurun_out/r2f_bench_train_n2.err
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --no-eager --no-cpu-baseline --comm-dtype f32 --torch-adam > gpurun_out/r2f_bench_train_n2_r1cfg.json 2> gpurun_out/r2f_bench_train_n2_r1cfg.err
timeout 200 python bench.py --no-eager --no-cpu-baseline > gpurun_out/r2f_bench_train.json 2> gpurun_out/r2f_bench_train.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2f_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        sm = d['stage_ms_per_step']
        print(f.split('r2f_bench_')[1][:-5], 'n', d['n_gpus'], 'ms', round(d['ms_per_step'], 2), 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), d['e2e']['ms_per_step'], 'cpu', (d.get('cpu_baseline') or {}).get('value'), d.get('collective'),
              {k: round(v, 2) for k, v in list(sm.items())[:10]})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-1200:])
PY
