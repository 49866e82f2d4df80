# round 2, call u: chunk-geometry sweep of the fused attention kernels (rows per chunk x buffers), per shape
mkdir -p gpurun_out
for cfg in "0:0" "128:1" "128:2" "64:1" "64:2" "32:1" "32:2" "16:2"; do
  R=${cfg%%:*}; NB=${cfg##*:}
  DGSCT_FA_R=$R DGSCT_FA_NBUF=$NB timeout 200 python scripts/stage_table_bwd.py > gpurun_out/r2u_stage_R${R}_nb$NB.log 2>&1
  echo "== R=$R nbuf=$NB"; grep -E "^stage|bwd_token|tok_update|x_update" gpurun_out/r2u_stage_R${R}_nb$NB.log
done
