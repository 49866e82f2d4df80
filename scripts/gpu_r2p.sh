# round 2, call p: bulk-copy issue dealt out over all warps + frame-shape-only splits -- tests, A/B, stage tables, kernel durations
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py tests/test_gpu_round2.py tests/test_gpu_zfullsize.py tests/test_gpu_fused_attn.py -m gpu -q --timeout 300 --tb=line 2>&1 | grep -v "UserWarning\|^  " | cut -c1-600 | tail -30 > gpurun_out/r2p_tests.log
tail -20 gpurun_out/r2p_tests.log
timeout 200 python bench.py --no-eager --no-cpu-baseline --steps 15 > gpurun_out/r2p_bench_train.json 2> gpurun_out/r2p_bench_train.err
DGSCT_NO_FUSED_ATTN=1 timeout 200 python bench.py --no-eager --no-cpu-baseline --steps 15 > gpurun_out/r2p_bench_train_chain.json 2> gpurun_out/r2p_bench_train_chain.err
timeout 200 python bench.py --mode fwd --no-eager --no-cpu-baseline > gpurun_out/r2p_bench_fwd.json 2> gpurun_out/r2p_bench_fwd.err
timeout 300 python scripts/stage_table_bwd.py > gpurun_out/r2p_stage_bwd.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2p_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        sm = d['stage_ms_per_step']
        print(f.split('r2p_bench_')[1][:-5], 'ms', round(d['ms_per_step'], 2), 'fwd', round((d.get('forward_only') or {}).get('ms_per_step', 0), 2), {k: round(v, 2) for k, v in list(sm.items())[:12]})
        print('   roofline', d['roofline']['stage'], round(d['roofline']['frac'], 3), {k: (round(v['achieved']), v['unit'], round(v['frac'], 3)) for k, v in d['roofline']['stages'].items()})
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-800:])
PY
grep -E "^stage|bwd_token|col_softmax|row_softmax|tok_update|x_update|TOTAL|weighted" gpurun_out/r2p_stage_bwd.log
for sh in vis_s0 aud_s2 vis_s2; do
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -k regex:"attn|dtp_prep" --csv --log-file /tmp/d_$sh.csv python scripts/one_call.py $sh 80 train > /dev/null 2>&1
  python - $sh <<'PY'
import csv, sys, re, collections
rows = list(csv.reader(l for l in open(f'/tmp/d_{sys.argv[1]}.csv') if l.startswith('"')))[1:]
d = collections.OrderedDict()
for r in rows:
    k = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("dgsct::fa::", "")
    d.setdefault((r[0], k), {})[r[-3]] = float(r[-1].replace(",", "")), r[-2]
print(sys.argv[1], " | ".join(f"{k[1]}: " + " ".join(f"{m.split('__')[1][:14]}={v[0]:.4g}{v[1]}" for m, v in ms.items()) for k, ms in d.items()))
PY
done
