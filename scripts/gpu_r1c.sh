# call C: full bench lines (fwd, train), train launch list, bwd stage table
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 python bench.py --mode fwd > gpurun_out/c_bench_fwd.json 2> gpurun_out/c_bench_fwd.err
timeout 600 python bench.py > gpurun_out/c_bench_train.json 2> gpurun_out/c_bench_train.err
timeout 300 python scripts/stage_table_bwd.py > gpurun_out/c_stage_bwd.log 2>&1
export DGSCT_BENCH_QUICK=1 DGSCT_BENCH_PROFILE=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/c_launches_train.csv python bench.py --steps 2 --warmup 3 > gpurun_out/c_prof2.log 2>&1
du -sh gpurun_out; tail -c 300 gpurun_out/c_prof2.log
