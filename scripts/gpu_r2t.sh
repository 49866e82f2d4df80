# round 2, call t: launch lists of one eager step (train, forward), ncu --set full of the fused attention kernels (final build), new tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_attn.py tests/test_gpu_backward.py -m gpu -q --timeout 300 --tb=short 2>&1 | tail -15 | cut -c1-300 > gpurun_out/r2t_tests.log
tail -5 gpurun_out/r2t_tests.log
DGSCT_BENCH_QUICK=1 DGSCT_BENCH_PROFILE=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2t_launches_train.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-eager > gpurun_out/r2t_ncu_train.log 2>&1
DGSCT_BENCH_QUICK=1 DGSCT_BENCH_PROFILE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2t_launches_fwd.csv python bench.py --mode fwd --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-eager > gpurun_out/r2t_ncu_fwd.log 2>&1
wc -l gpurun_out/r2t_launches_train.csv gpurun_out/r2t_launches_fwd.csv
for sh in vis_s0 aud_s2 vis_s2; do
  timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:"attn|dtp_prep|vec_linear" -o /tmp/r2t_full_$sh python scripts/one_call.py $sh 80 train > gpurun_out/r2t_ncu_$sh.log 2>&1
  ncu -i /tmp/r2t_full_$sh.ncu-rep --page raw --csv > /tmp/r2t_full_${sh}_raw.csv 2>/dev/null
  echo "== $sh (one adapter call, BT=80: forward with kept activations + backward)" >> gpurun_out/r2t_full_fused_attention_summary.txt
  python scripts/ncu_summary.py /tmp/r2t_full_${sh}_raw.csv >> gpurun_out/r2t_full_fused_attention_summary.txt 2>&1
done
cut -c1-260 gpurun_out/r2t_full_fused_attention_summary.txt
