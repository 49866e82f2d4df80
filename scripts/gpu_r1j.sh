# call J (final of the round): tests, bench lines, ncu launch lists of one eager step (fwd, train), pipeline probe
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -6 > gpurun_out/j_tests.log
tail -2 gpurun_out/j_tests.log
timeout 300 python bench.py > gpurun_out/j_bench_train.json 2> gpurun_out/j_bench_train.err
timeout 200 python bench.py --mode fwd > gpurun_out/j_bench_fwd.json 2> gpurun_out/j_bench_fwd.err
export DGSCT_BENCH_QUICK=1 DGSCT_BENCH_PROFILE=1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/j_launches_fwd.csv python bench.py --mode fwd --steps 2 --warmup 3 > gpurun_out/j_prof1.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/j_launches_train.csv python bench.py --steps 2 --warmup 3 > gpurun_out/j_prof2.log 2>&1
unset DGSCT_BENCH_QUICK DGSCT_BENCH_PROFILE
timeout 120 python scripts/gemm_pipe_probe.py > gpurun_out/j_pipe_probe.log 2>&1
du -sh gpurun_out; tail -2 gpurun_out/j_prof2.log | cut -c1-200
