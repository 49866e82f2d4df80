# round-1 re-entry call A: gpu tests, both bench modes, reference arm, ncu launch lists + full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.log
( time timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 ) 2>&1 | tail -15 > gpurun_out/a_tests.log
tail -4 gpurun_out/a_tests.log
timeout 600 python bench.py --mode fwd > gpurun_out/a_bench_fwd.json 2> gpurun_out/a_bench_fwd.err
timeout 600 python bench.py > gpurun_out/a_bench_train.json 2> gpurun_out/a_bench_train.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err
export DGSCT_BENCH_QUICK=1 DGSCT_BENCH_PROFILE=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/a_launches_fwd.csv python bench.py --mode fwd --steps 2 --warmup 3 > gpurun_out/a_prof1.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/a_launches_train.csv python bench.py --steps 2 --warmup 3 > gpurun_out/a_prof2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernel -c 28 -o gpurun_out/a_full_gemm python bench.py --mode fwd --steps 2 --warmup 3 > gpurun_out/a_prof3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"ln|softmax|scale_cols|spatial_map|bn_" -c 16 -o gpurun_out/a_full_row python bench.py --mode fwd --steps 2 --warmup 3 > gpurun_out/a_prof4.log 2>&1
ls -la gpurun_out
for f in fwd train ref; do tail -c 1500 gpurun_out/a_bench_$f.json; echo; tail -c 400 gpurun_out/a_bench_$f.err; done
