"""Where does the bf16 error come from?  CPU experiment with the ORACLE (test infrastructure): round ONE tensor at a time to
bf16 -- in the forward pass (value) or in the backward pass (its gradient) -- and measure the rel-L2 change of the adapter
output and of the gradients against the unrounded fp32 oracle.  Also measures PyTorch's own bf16 autocast of the same ops.
Output committed as profiles/r02_bf16_error_budget.txt; the fixed bf16 bars of tests/test_gpu_backward.py and
tests/test_gpu_round2.py are read off these tables.   usage: python scripts/bf16_error_budget.py > profiles/r02_bf16_error_budget.txt"""
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dgsct_oracle as O   # noqa: E402


class RB(torch.autograd.Function):
    """identity that rounds the value (forward) and / or the gradient (backward) to bf16"""

    @staticmethod
    def forward(ctx, t, f, b):
        ctx.b = b
        return t.bfloat16().float() if f else t.clone()

    @staticmethod
    def backward(ctx, g):
        return (g.bfloat16().float() if ctx.b else g), None, None


def fwd(cfg, p, x, other, R):
    def r(t, name):
        return RB.apply(t, ("f:" + name) in R or "F" in R, ("b:" + name) in R or "B" in R)

    def w(n):
        return RB.apply(p[n], "W" in R or ("W:" + n) in R, False)
    if cfg.resampler == "conv":
        wc = w("conv_adapter.weight").reshape(cfg.N, cfg.No)
        u = r(torch.einsum("nm,bmc->bnc", wc, other) + p["conv_adapter.bias"][None, :, None], "U")
        V = u @ w("fc.weight").t() + p["fc.bias"]
    else:   # bicubic stencil with fp32 coefficients (the engine's sparse resampler)
        M = O.bicubic_matrix(int(math.sqrt(cfg.No)), int(math.sqrt(cfg.N))).float()
        u = r(torch.einsum("nm,bmc->bnc", M, other), "U")
        V = u @ w("fc.weight").t() + p["fc.bias"]
    obar = V.mean(1)
    V = r(V, "V")
    tok, tokb = p["my_tokens"], w("my_tokens")
    P = r(torch.softmax(torch.einsum("kc,bnc->bkn", tokb, V), -1), "P")
    Tp = tok[None] + torch.einsum("bkn,bnc->bkc", P, V)
    Tpb = r(Tp, "Tp")
    Q = r(torch.softmax(torch.einsum("bnc,bkc->bnk", x, Tpb), -1), "Q")
    X1 = r(x + p["gate_av"] * torch.einsum("bnk,bkc->bnc", Q, Tpb), "X1")
    aq1 = torch.relu(obar @ w("fc_affine_audio_1.weight").t() + p["fc_affine_audio_1.bias"])
    vq1 = torch.relu(r(X1 @ w("fc_affine_video_1.weight").t() + p["fc_affine_video_1.bias"], "z1"))
    q = (aq1[:, None, :] * vq1).mean(1)
    h = torch.relu(q @ w("fc_affine_bottleneck.weight").t() + p["fc_affine_bottleneck.bias"])
    m = torch.sigmoid(h @ w("fc_affine_v_c_att.weight").t() + p["fc_affine_v_c_att.bias"])
    Xc = r(X1 * (1 + m[:, None, :]), "Xc")
    vq2 = torch.relu(r(Xc @ w("fc_affine_video_2.weight").t() + p["fc_affine_video_2.bias"], "z2"))
    aq2 = torch.relu(obar @ w("fc_affine_audio_2.weight").t() + p["fc_affine_audio_2.bias"])
    s = (vq2 * aq2[:, None, :]) @ p["fc_affine_v_s_att.weight"].reshape(-1) + p["fc_affine_v_s_att.bias"]
    sig, smap = torch.sigmoid(s), torch.softmax(torch.tanh(s), -1)
    Y = X1 * (cfg.alpha * m[:, None, :] + cfg.beta * sig[:, :, None] + (1 - cfg.alpha))
    if cfg.ln_before:
        Y = F.layer_norm(Y, (cfg.C,), p["ln_before.weight"], p["ln_before.bias"], cfg.eps_ln)
    Y = r(Y, "Y")
    ns = {}
    Z = O._grouped_pointwise(Y, w("down_sampler.weight"), cfg.g)
    if cfg.use_bn:
        Z = O._batchnorm(Z, p, "bn1", cfg, True, ns)
    U = r(O._grouped_pointwise(r(torch.relu(Z), "Zr"), w("up_sampler.weight"), cfg.g), "Uo")
    if cfg.use_bn:
        U = O._batchnorm(U, p, "bn2", cfg, True, ns)
    out = U
    if cfg.gate_before_ln:
        if cfg.use_gate:
            out = p["gate"] * out
        out = F.layer_norm(out, (cfg.C,), p["ln_post.weight"], p["ln_post.bias"], cfg.eps_ln)
    else:
        out = F.layer_norm(out, (cfg.C,), p["ln_post.weight"], p["ln_post.bias"], cfg.eps_ln)
        if cfg.use_gate:
            out = p["gate"] * out
    return out, smap


def rel(a, b):
    return float((a - b).norm() / b.norm())


PTS = ["U", "V", "P", "Tp", "Q", "X1", "z1", "Xc", "z2", "Y", "Zr", "Uo"]
KEYS = ["out", "x", "other", "fc.weight", "my_tokens", "fc_affine_video_1.weight", "down_sampler.weight", "fc.bias", "gate_av"]


def table(variant, shape, kw):
    C, N, Co, No = shape
    cfg = O.make_config(variant, C, N, Co, No, **kw)
    params = O.random_params(cfg, seed=21)
    x, other = O.random_inputs(cfg, 10, seed=22)
    g = torch.Generator().manual_seed(23)
    R_out = torch.randn(10, N, C, generator=g).bfloat16().float()
    R_smap = torch.randn(10, N, generator=g).bfloat16().float()
    xq, oq = x.bfloat16().float(), other.bfloat16().float()

    def grads(R):
        p = {k: v.clone().requires_grad_("running" not in k) for k, v in params.items()}
        xs, os_ = xq.clone().requires_grad_(True), oq.clone().requires_grad_(True)
        out, smap = fwd(cfg, p, xs, os_, R)
        ((out * R_out).sum() + (smap * R_smap).sum()).backward()
        d = {k: v.grad for k, v in p.items() if v.grad is not None}
        d["x"], d["other"], d["out"] = xs.grad, os_.grad, out.detach()
        return d

    ref = grads(set())
    keys = [k for k in KEYS if k in ref]
    print(f"\n== {variant} C={C} N={N} C'={Co} N'={No}, BT=10, train mode: rel-L2 against the unrounded fp32 oracle")
    print(" " * 16 + " ".join(f"{k[:10]:>10s}" for k in keys))

    def show(name, R):
        d = grads(R)
        print(f"{name:16s}" + " ".join(f"{rel(d[k], ref[k]):10.1e}" for k in keys))

    show("everything bf16", {"F", "B", "W"})
    show("weights only", {"W"})
    show("all fwd values", {"F"})
    show("all bwd grads", {"B"})
    for pt in PTS:
        show("fwd " + pt, {"f:" + pt})
    for pt in ("V", "Tp", "X1", "Y", "Uo"):
        show("bwd d" + pt, {"b:" + pt})
    # torch's own autocast of the oracle ops
    p = {k: v.clone().requires_grad_("running" not in k) for k, v in params.items()}
    xs, os_ = xq.clone().requires_grad_(True), oq.clone().requires_grad_(True)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        out, smap, _, _ = O.adapter_forward(cfg, p, xs, os_, training=True)
    ((out.float() * R_out).sum() + (smap.float().reshape(R_smap.shape) * R_smap).sum()).backward()
    d = {k: v.grad for k, v in p.items() if v.grad is not None}
    d["x"], d["other"], d["out"] = xs.grad, os_.grad, out.detach().float()
    print(f"{'torch autocast':16s}" + " ".join(f"{rel(d[k].float(), ref[k]):10.1e}" for k in keys))


if __name__ == "__main__":
    print(__doc__)
    table("ave", (768, 144, 384, 256), {})
    table("ave", (384, 576, 192, 1024), {})
    table("s4", (1536, 36, 768, 64), dict(use_bn=False, use_gate=False))
    table("s4", (768, 144, 384, 256), dict(use_bn=False, use_gate=False))
    table("avqa", (768, 144, 384, 256), dict(K=2, g=4, use_bn=False, use_gate=False))
