# round 2, call x: where a SMALL tcgen05 GEMM launch spends its time (per-SASS stall samples of three launches of a deep-stage call)
mkdir -p gpurun_out
for sk in 3 13 21; do
  timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:gemm_tc_kernel -s $sk -c 1 -o /tmp/gemm_$sk python scripts/one_call.py aud_s2 80 train > gpurun_out/r2x_ncu_$sk.log 2>&1
  ncu -i /tmp/gemm_$sk.ncu-rep --page source --csv --print-source sass > /tmp/gemm_$sk.sass.csv 2>/dev/null
  gzip -c /tmp/gemm_$sk.sass.csv > gpurun_out/r2x_gemm_$sk.sass.csv.gz
  ncu -i /tmp/gemm_$sk.ncu-rep --page raw --csv > /tmp/gemm_$sk.raw.csv 2>/dev/null
  echo "== launch $sk"; python scripts/ncu_summary.py /tmp/gemm_$sk.raw.csv | cut -c1-260
  python scripts/ncu_source_hot.py /tmp/gemm_$sk.sass.csv 30 | cut -c1-200
done
