"""Fused latent-token attention (attn_fused.cuh) against the six-launch GEMM chain on the same inputs: forward outputs,
input gradients and every parameter gradient of one adapter call, per shape (rel-L2 fused vs chain; both are bf16 paths, so
the differences are rounding noise of a few 1e-3, not 1e-1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dataclasses
import torch
from dgsct_b200.adapter import adapter_backward_raw, adapter_forward_raw
from oracle import dgsct_oracle as O


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-30))


def run(cfg, p, x, o, do, dm, fused):
    if fused:
        os.environ.pop("DGSCT_NO_FUSED_ATTN", None)
        os.environ.pop("DGSCT_NO_FUSED_ATTN_BWD", None)
    else:
        os.environ["DGSCT_NO_FUSED_ATTN"] = "1"
    cd = dataclasses.asdict(cfg)
    keep = []
    out, smap, _ = adapter_forward_raw(cd, p, x, o, training=True, keep=keep)
    dx, dother, grads = adapter_backward_raw(cd, p, x, o, do, dm, None, training=True, saved=keep[0])
    torch.cuda.synchronize()
    return out, smap, dx, dother, grads


cases = [("ave", 192, 2304, 96, 4096, 80, {}), ("ave", 96, 4096, 192, 2304, 80, {}), ("ave", 384, 576, 192, 1024, 80, {}),
         ("ave", 192, 1024, 384, 576, 80, {}), ("ave", 768, 144, 384, 256, 80, {}), ("ave", 384, 256, 768, 144, 80, {}),
         ("ave", 1536, 36, 768, 64, 80, {}), ("ave", 768, 64, 1536, 36, 80, {}),
         ("s4", 96, 4096, 192, 2304, 20, dict(use_bn=False, use_gate=False)),
         ("avqa", 768, 144, 384, 256, 60, dict(K=2, g=4, use_bn=False, use_gate=False)),
         ("ave", 320, 100, 128, 49, 7, {}), ("ave", 1024, 49, 512, 64, 6, {})]
for variant, C, N, Co, No, BT, kw in cases:
    cfg = O.make_config(variant, C, N, Co, No, **kw)
    params = O.random_params(cfg, seed=11)
    x, other = O.random_inputs(cfg, BT, seed=12)
    g = torch.Generator().manual_seed(13)
    do = torch.randn(BT, N, C, generator=g).bfloat16().cuda()
    dm = torch.randn(BT, N, generator=g).bfloat16().cuda()
    p = {k: v.float().cuda().contiguous() for k, v in params.items()}
    xq, oq = x.bfloat16().cuda(), other.bfloat16().cuda()
    A = run(cfg, p, xq, oq, do, dm, True)
    B = run(cfg, p, xq, oq, do, dm, False)
    errs = {"out": rel(A[0], B[0]), "smap": rel(A[1], B[1]), "dx": rel(A[2], B[2]), "dother": rel(A[3], B[3])}
    for k in B[4]:
        if float(B[4][k].abs().max()) > 1e-9:
            errs[k] = rel(A[4][k], B[4][k])
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    print(f"{variant} C={C} N={N} BT={BT}: " + " ".join(f"{k}={v:.1e}" for k, v in errs.items() if k in ("out", "smap", "dx", "dother", "my_tokens"))
          + " | worst: " + " ".join(f"{k}={v:.1e}" for k, v in worst))
