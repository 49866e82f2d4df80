# round 2, call ae: rows per split of the forward token kernel (partials vs chain length), per shape
mkdir -p gpurun_out
for r in 256 384 576 768 1152 2304; do
  DGSCT_FA_SPLIT_ROWS=$r timeout 200 python scripts/stage_table.py > gpurun_out/r2ae_stage_$r.log 2>&1
  echo "== rows per split $r"; grep -E "^stage|tok_update" gpurun_out/r2ae_stage_$r.log
done
