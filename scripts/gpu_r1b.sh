# call B: gpu tests log, fwd launch list (one eager step), full captures exported to CSV on the box (reports are too big to ship)
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -6 > gpurun_out/b_tests.log
cat gpurun_out/b_tests.log
export DGSCT_BENCH_QUICK=1 DGSCT_BENCH_PROFILE=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/b_launches_fwd.csv python bench.py --mode fwd --steps 2 --warmup 3 > gpurun_out/b_prof1.log 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:gemm_tc_kernel -c 14 -o /tmp/full_gemm python bench.py --mode fwd --steps 2 --warmup 3 > gpurun_out/b_prof3.log 2>&1
ncu -i /tmp/full_gemm.ncu-rep --page raw --csv > gpurun_out/b_full_gemm_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:"ln|softmax|scale_cols|spatial_map|channel_q|scale_to" -c 8 -o /tmp/full_row python bench.py --mode fwd --steps 2 --warmup 3 > gpurun_out/b_prof4.log 2>&1
ncu -i /tmp/full_row.ncu-rep --page raw --csv > gpurun_out/b_full_row_raw.csv 2>/dev/null
ls -la /tmp/*.ncu-rep gpurun_out
du -sh gpurun_out
