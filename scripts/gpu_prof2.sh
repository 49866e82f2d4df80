mkdir -p gpurun_out
export DGSCT_BENCH_QUICK=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|out_ln|gate_ln" -s 0 -c 12 -o gpurun_out/prof_s0 python bench.py --steps 1 --warmup 3 > gpurun_out/prof_s0.log 2>&1
tail -3 gpurun_out/prof_s0.log
