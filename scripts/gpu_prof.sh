mkdir -p gpurun_out
# launch list: every kernel with its device time (first profiled step only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/prof_bench.log 2>&1
tail -c 600 gpurun_out/prof_bench.log
# full capture of the first tensor-core GEMMs of a step (aud s0 resample conv + fc ...)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 700 -c 6 -o gpurun_out/prof_tc python bench.py --steps 2 --warmup 3 > gpurun_out/prof_bench2.log 2>&1
ls -la gpurun_out/
