# round 2, call ac (2 GPUs): final build at N=2 (train step under torchrun, reference arm), DataParallel / DDP tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q --timeout 300 --tb=short 2>&1 | tail -6 | cut -c1-300 > gpurun_out/r2ac_tests.log
tail -3 gpurun_out/r2ac_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --no-eager > gpurun_out/r2ac_bench_train_n2.json 2> gpurun_out/r2ac_bench_train_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/r2ac_bench_ref_n2.json 2> gpurun_out/r2ac_bench_ref_n2.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2ac_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('r2ac_bench_')[1][:-5], 'n', d['n_gpus'], 'ms', round(d['ms_per_step'], 2), 'value', round(d['value'], 1), 'e2e', {k: v for k, v in d['e2e'].items() if k in ('value', 'ms_per_step')}, 'cpu', (d.get('cpu_baseline') or {}).get('value'), d.get('collective'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-1500:])
PY
