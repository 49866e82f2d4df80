# round 2, call m: ncu --set full of the fused attention kernels at three shapes; slice-invariance numbers
mkdir -p gpurun_out
timeout 300 python scripts/slice_check.py > gpurun_out/r2m_slice.log 2>&1; cat gpurun_out/r2m_slice.log | cut -c1-400
for sh in vis_s0 aud_s2 vis_s3 vis_s2; do
  timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:"attn|dtp_prep" -o gpurun_out/r2m_full_$sh python scripts/one_call.py $sh 80 train > gpurun_out/r2m_ncu_$sh.log 2>&1
  ncu -i gpurun_out/r2m_full_$sh.ncu-rep --page raw --csv > gpurun_out/r2m_full_${sh}_raw.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/r2m_full_${sh}_raw.csv > gpurun_out/r2m_full_${sh}_summary.txt 2>&1
  echo "== $sh"; cat gpurun_out/r2m_full_${sh}_summary.txt
done
ls -la gpurun_out/*.ncu-rep
