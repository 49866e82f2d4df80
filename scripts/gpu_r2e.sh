# round 2, call e (2 GPUs): full GPU suite incl. the nn.DataParallel test, bench at N=1 and N=2 (bf16 wire + flat Adam vs fp32 + torch Adam)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --tb=short 2>&1 | tail -80 > gpurun_out/r2e_tests.log
tail -8 gpurun_out/r2e_tests.log
timeout 300 python bench.py --no-eager > gpurun_out/r2e_bench_train.json 2> gpurun_out/r2e_bench_train.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --no-eager > gpurun_out/r2e_bench_train_n2.json 2> gpurun_out/r2e_bench_train_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --no-eager --no-cpu-baseline --comm-dtype f32 --torch-adam > gpurun_out/r2e_bench_train_n2_r1cfg.json 2> gpurun_out/r2e_bench_train_n2_r1cfg.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/r2e_bench_ref_n2.json 2> gpurun_out/r2e_bench_ref_n2.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2e_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('r2e_bench_')[1][:-5], 'n', d['n_gpus'], 'ms', round(d['ms_per_step'], 2), 'value', round(d['value'], 1), 'e2e', d['e2e'], 'cpu', d.get('cpu_baseline'), d.get('collective'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-1500:])
PY
