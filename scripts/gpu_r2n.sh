# round 2, call n: where the fused attention kernels wait -- per-source-line stall samples (ncu --import-source); determinism probe
mkdir -p gpurun_out
for spec in "vis_s0:tok_attn_fwd" "vis_s0:x_attn_fwd_kernel" "vis_s0:x_attn_bwd" "aud_s2:tok_attn_fwd" "aud_s2:x_attn_fwd_kernel"; do
  sh=${spec%%:*}; k=${spec##*:}
  timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"$k" -c 1 -o /tmp/src_${sh}_$k python scripts/one_call.py $sh 80 train > gpurun_out/r2n_ncu_${sh}_$k.log 2>&1
  ncu -i /tmp/src_${sh}_$k.ncu-rep --page source --csv --print-source cuda > /tmp/src_${sh}_$k.csv 2>/dev/null
  echo "== $sh $k"; python scripts/ncu_source_hot.py /tmp/src_${sh}_$k.csv 45 2>&1 | tee gpurun_out/r2n_hot_${sh}_$k.txt | cut -c1-200
  ncu -i /tmp/src_${sh}_$k.ncu-rep --page raw --csv > /tmp/raw.csv 2>/dev/null; python scripts/ncu_summary.py /tmp/raw.csv | cut -c1-260
done
python - <<'PY'
import os, sys, torch
sys.path.insert(0, '.')
from tests.test_gpu_zfullsize import _inputs, SHAPES, BT, to_dev, cfg_dict
from dgsct_b200.adapter import adapter_backward_raw
for shape in ("vis_s1", "aud_s2"):
    cfg, params, x, other = _inputs(shape, 61)
    C, N, Co, No = SHAPES[shape]
    p = to_dev(params); xd, od = x.cuda(), other.cuda()
    g = torch.Generator().manual_seed(62)
    d_out = (torch.randn(BT, N, C, generator=g) * 1e-1).bfloat16().cuda(); d_smap = torch.randn(BT, N, generator=g).bfloat16().cuda()
    a = adapter_backward_raw(cfg_dict(cfg), p, xd, od, d_out, d_smap, None, training=False)[0]
    for i in range(3):
        b = adapter_backward_raw(cfg_dict(cfg), p, xd, od, d_out, d_smap, None, training=False)[0]
        diff = (a != b)
        idx = diff.nonzero()
        print(shape, "differing dx elements:", int(diff.sum()), "frames", sorted(set(idx[:, 0].tolist()))[:10], "rows", sorted(set(idx[:, 1].tolist()))[:12], "cols", sorted(set(idx[:, 2].tolist()))[:12],
              "max abs", float((a.float() - b.float()).abs().max()))
PY
