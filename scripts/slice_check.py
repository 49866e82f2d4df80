"""Full-batch backward vs per-clip backward (tests/test_gpu_zfullsize.py) with the fused attention kernels on / off: rel-L2 of
dx / dother per clip."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_zfullsize import _inputs, SHAPES, BT, CLIP, to_dev, cfg_dict
from tests.golden_util import rel_l2
from dgsct_b200.adapter import adapter_backward_raw

for mode in ("fused", "fused_fwd_only", "chain"):
    os.environ.pop("DGSCT_NO_FUSED_ATTN", None); os.environ.pop("DGSCT_NO_FUSED_ATTN_BWD", None)
    if mode == "chain": os.environ["DGSCT_NO_FUSED_ATTN"] = "1"
    if mode == "fused_fwd_only": os.environ["DGSCT_NO_FUSED_ATTN_BWD"] = "1"
    for shape in ("vis_s1", "aud_s2"):
        cfg, params, x, other = _inputs(shape, 61)
        C, N, Co, No = SHAPES[shape]
        p = to_dev(params)
        xd, od = x.cuda(), other.cuda()
        g = torch.Generator().manual_seed(62)
        d_out = (torch.randn(BT, N, C, generator=g) * 1e-1).bfloat16().cuda()
        d_smap = torch.randn(BT, N, generator=g).bfloat16().cuda()
        dx, do, grads = adapter_backward_raw(cfg_dict(cfg), p, xd, od, d_out, d_smap, None, training=False)
        dxr, dor, _ = adapter_backward_raw(cfg_dict(cfg), p, xd, od, d_out, d_smap, None, training=False)
        line = [f"repeat dx={rel_l2(dx.float().cpu(), dxr.float().cpu()):.1e} do={rel_l2(do.float().cpu(), dor.float().cpu()):.1e}"]
        for lo in range(0, BT, CLIP):
            sl = slice(lo, lo + CLIP)
            dx2, do2, _ = adapter_backward_raw(cfg_dict(cfg), p, xd[sl].contiguous(), od[sl].contiguous(), d_out[sl].contiguous(),
                                               d_smap[sl].contiguous(), None, training=False)
            line.append(f"{lo}:{rel_l2(dx[sl].float().cpu(), dx2.float().cpu()):.1e}/{rel_l2(do[sl].float().cpu(), do2.float().cpu()):.1e}")
        print(mode, shape, " ".join(line))
