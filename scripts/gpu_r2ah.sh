# round 2, call ah: DRAM traffic of ONE adapter call, every kernel (ncu --set full): stage-0 audio forward (round 1's 6.2x case)
# and a stage-0 visual forward + backward
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --profile-from-start off -o /tmp/full_aud_s0_fwd python scripts/one_call.py aud_s0 80 fwd > gpurun_out/r2ah_ncu_aud_s0_fwd.log 2>&1
ncu -i /tmp/full_aud_s0_fwd.ncu-rep --page raw --csv > /tmp/full_aud_s0_fwd.csv 2>/dev/null
python scripts/ncu_summary.py /tmp/full_aud_s0_fwd.csv | cut -c1-200 > gpurun_out/r2ah_full_call_aud_s0_fwd_summary.txt
timeout 700 ncu --set full --clock-control none --profile-from-start off -o /tmp/full_vis_s0 python scripts/one_call.py vis_s0 80 train > gpurun_out/r2ah_ncu_vis_s0.log 2>&1
ncu -i /tmp/full_vis_s0.ncu-rep --page raw --csv > /tmp/full_vis_s0.csv 2>/dev/null
python scripts/ncu_summary.py /tmp/full_vis_s0.csv | cut -c1-200 > gpurun_out/r2ah_full_call_vis_s0_summary.txt
python - <<'PY'
for n in ('aud_s0_fwd', 'vis_s0'):
    us = rd = wr = 0.0; k = 0
    for l in open(f'gpurun_out/r2ah_full_call_{n}_summary.txt').read().splitlines()[1:]:
        p = l[64:].split()
        try:
            us += float(p[0]); rd += float(p[1]); wr += float(p[2]); k += 1
        except Exception:
            pass
    print(n, 'kernels', k, 'sum us', round(us, 1), 'DRAM read MB', round(rd, 1), 'write MB', round(wr, 1))
PY
