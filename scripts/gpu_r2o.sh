# round 2, call o: determinism probe; source-level stall samples brought back as (gzipped) CSV
mkdir -p gpurun_out
REP=6 timeout 600 python scripts/determinism_probe.py 2>&1 | tee gpurun_out/r2o_determinism.log | cut -c1-300
for spec in "vis_s0:tok_attn_fwd" "vis_s0:x_attn_fwd_kernel" "aud_s2:tok_attn_fwd" "aud_s2:x_attn_bwd"; do
  sh=${spec%%:*}; k=${spec##*:}
  timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"$k" -c 1 -o /tmp/src_${sh}_$k python scripts/one_call.py $sh 80 train > gpurun_out/r2o_ncu_${sh}_$k.log 2>&1
  ncu -i /tmp/src_${sh}_$k.ncu-rep --page source --csv --print-source cuda > /tmp/src_${sh}_$k.cuda.csv 2>/tmp/err1.txt
  ncu -i /tmp/src_${sh}_$k.ncu-rep --page source --csv --print-source sass > /tmp/src_${sh}_$k.sass.csv 2>/tmp/err2.txt
  ls -la /tmp/src_${sh}_$k.*; head -c 600 /tmp/src_${sh}_$k.cuda.csv; echo; head -c 300 /tmp/err1.txt
  gzip -c /tmp/src_${sh}_$k.cuda.csv > gpurun_out/r2o_src_${sh}_$k.cuda.csv.gz
  gzip -c /tmp/src_${sh}_$k.sass.csv > gpurun_out/r2o_src_${sh}_$k.sass.csv.gz
done
ls -la gpurun_out/r2o_*
