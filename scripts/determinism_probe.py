"""Run-to-run repeatability of one adapter call: which forward taps / backward outputs change bits between identical calls."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_zfullsize import _inputs, SHAPES, BT, to_dev, cfg_dict
from dgsct_b200.adapter import adapter_backward_raw, adapter_forward_raw

REP = int(os.environ.get("REP", "8"))
for mode in ("fused", "fused_fwd_only", "chain"):
    os.environ.pop("DGSCT_NO_FUSED_ATTN", None); os.environ.pop("DGSCT_NO_FUSED_ATTN_BWD", None)
    if mode == "chain": os.environ["DGSCT_NO_FUSED_ATTN"] = "1"
    if mode == "fused_fwd_only": os.environ["DGSCT_NO_FUSED_ATTN_BWD"] = "1"
    for shape in ("vis_s1", "aud_s2", "vis_s0"):
        cfg, params, x, other = _inputs(shape, 61)
        C, N, Co, No = SHAPES[shape]
        p = to_dev(params); xd, od = x.cuda(), other.cuda()
        g = torch.Generator().manual_seed(62)
        d_out = (torch.randn(BT, N, C, generator=g) * 1e-1).bfloat16().cuda(); d_smap = torch.randn(BT, N, generator=g).bfloat16().cuda()
        cd = cfg_dict(cfg)
        ref = None; fdiff = {}
        for i in range(REP):
            keep = []
            out, smap, _, taps = adapter_forward_raw(cd, p, xd, od, training=False, want_taps=True, keep=keep)
            cur = dict(taps, out=out, smap=smap)
            if ref is None: ref = {k: v.clone() for k, v in cur.items()}; saved = keep[0]
            else:
                for k, v in cur.items():
                    if not torch.equal(v, ref[k]): fdiff[k] = fdiff.get(k, 0) + 1
        bref = None; bdiff = {}
        for i in range(REP):
            dx, do, grads = adapter_backward_raw(cd, p, xd, od, d_out, d_smap, None, training=False, saved=saved)
            cur = dict(grads, dx=dx, dother=do)
            if bref is None: bref = {k: v.clone() for k, v in cur.items()}
            else:
                for k, v in cur.items():
                    if not torch.equal(v, bref[k]):
                        bdiff[k] = bdiff.get(k, 0) + 1
        print(mode, shape, "fwd changed:", fdiff, "| bwd changed:", {k: bdiff[k] for k in ("dx", "dother") if k in bdiff}, "params:", sorted(k for k in bdiff if k not in ("dx", "dother"))[:30])
