"""CPU oracle for the DG-SCT cross-modal adapter hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch restatement (token-major, explicit matmuls, dtype-generic torch on
the CPU) of ``VisualAdapter.forward`` of haoyi-duan/DG-SCT.  It is the checker for the CUDA
engine in ``dgsct_b200/`` and the CPU arm of ``bench.py``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
it; the product path (``dgsct_b200``) never does.

Parity pinning: the reference ships no tests or golden vectors.  The oracle is
pinned against outputs of the reference's OWN classes, imported read-only from /root/reference by
``tests/golden/make_golden.py`` (one subprocess per task tree) and committed as fixtures under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them in fp64 (forward, both return
values, BN running-stat updates, and every input/parameter gradient).

Reference lines restated (all relative to /root/reference):
  resampler (conv)     DG-SCT/AVE/nets/net_trans.py:553-555
  resampler (bicubic)  DG-SCT/AVS/avs_scripts/avs_s4/model/PVT_AVSModel.py:190-197
  latent-token attn    DG-SCT/AVE/nets/net_trans.py:572-589
  channel gate         DG-SCT/AVE/nets/net_trans.py:592-598
  spatial gate         DG-SCT/AVE/nets/net_trans.py:602-609
  temporal gate        few-shot/nets/net_trans.py:534-535
  gate combine         DG-SCT/AVE/nets/net_trans.py:611-613 ; few-shot/nets/net_trans.py:537-549
  bottleneck           DG-SCT/AVE/nets/net_trans.py:626-643
  post LN / gate       DG-SCT/AVE/nets/net_trans.py:667-672 ; S4 order PVT_AVSModel.py:308-313

Layout convention here: activations are token-major ``(BT, N, C)``; the reference's call boundary
``(BT, C, N, 1)`` is a permuted view of exactly that buffer (net_trans.py:891).
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclasses.dataclass
class AdapterConfig:
    """Static description of one adapter instance (mirrors the constructor args + variant flags)."""
    C: int                 # width of the adapted stream   (input_dim == output_dim == linear_out)
    N: int                 # tokens of the adapted stream  (conv_dim_out)
    Co: int                # width of the other stream     (linear_in)
    No: int                # tokens of the other stream    (conv_dim_in)
    K: int = 32            # latent tokens (num_tk / opt.num_tokens)
    r: int = 8             # reduction_factor
    g: int = 2             # opt.num_conv_group
    T: int = 10            # frames per clip (temporal gate only)
    alpha: float = 0.3
    beta: float = 0.05
    gamma: float = 0.0     # 0 => temporal gate does not enter the output
    resampler: str = "conv"        # "conv" (conv_adapter+fc) | "bicubic" (fc + F.interpolate)
    ln_before: bool = True
    ln_post: bool = True
    gate_before_ln: bool = False   # S4/MS3: gate * out THEN ln_post
    use_bn: bool = True
    use_gate: bool = True
    temporal: bool = False         # compute + return temporal_att_maps (3-tuple variants)
    eps_ln: float = 1e-5
    eps_bn: float = 1e-5
    bn_momentum: float = 0.1

    @property
    def Cd(self) -> int:           # bottleneck width
        return self.C // self.r


# Variant presets.  Dimensions are supplied by the caller.
VARIANTS = {
    "ave":      dict(alpha=0.3, beta=0.05, gamma=0.0, resampler="conv", ln_before=True, ln_post=True,
                     gate_before_ln=False, temporal=False, T=10),
    "avvp":     dict(alpha=0.3, beta=0.05, gamma=0.0, resampler="conv", ln_before=True, ln_post=True,
                     gate_before_ln=False, temporal=False, T=10),
    "avqa":     dict(alpha=0.3, beta=0.05, gamma=0.0, resampler="conv", ln_before=True, ln_post=True,
                     gate_before_ln=False, temporal=False, T=10),
    "s4":       dict(alpha=0.3, beta=0.05, gamma=0.0, resampler="bicubic", ln_before=False, ln_post=True,
                     gate_before_ln=True, temporal=False, T=5),
    "ms3":      dict(alpha=0.2, beta=0.1, gamma=0.0, resampler="conv", ln_before=False, ln_post=True,
                     gate_before_ln=True, temporal=False, T=5),
    "pretrain": dict(alpha=0.3, beta=0.01, gamma=0.05, resampler="conv", ln_before=True, ln_post=True,
                     gate_before_ln=False, temporal=True, T=10),
    "fewshot":  dict(alpha=1.0, beta=1.0, gamma=1.0, resampler="conv", ln_before=True, ln_post=True,
                     gate_before_ln=False, temporal=True, T=10),
}


def make_config(variant: str, C: int, N: int, Co: int, No: int, **kw) -> AdapterConfig:
    base = dict(VARIANTS[variant])
    base.update(kw)
    return AdapterConfig(C=C, N=N, Co=Co, No=No, **base)


def param_shapes(cfg: AdapterConfig) -> Dict[str, Tuple[int, ...]]:
    """state_dict names/shapes of the reference module, parameters + buffers."""
    C, N, Co, No, K, Cd, g = cfg.C, cfg.N, cfg.Co, cfg.No, cfg.K, cfg.Cd, cfg.g
    H = C // 2
    s: Dict[str, Tuple[int, ...]] = {
        "conv_adapter.weight": (N, No, 1, 1), "conv_adapter.bias": (N,),
        "fc.weight": (C, Co), "fc.bias": (C,),
        "fc_affine_audio_1.weight": (C, C), "fc_affine_audio_1.bias": (C,),
        "fc_affine_video_1.weight": (C, C), "fc_affine_video_1.bias": (C,),
        "fc_affine_bottleneck.weight": (H, C), "fc_affine_bottleneck.bias": (H,),
        "fc_affine_video_2.weight": (H, C), "fc_affine_video_2.bias": (H,),
        "fc_affine_audio_2.weight": (H, C), "fc_affine_audio_2.bias": (H,),
        "fc_affine_v_s_att.weight": (1, H), "fc_affine_v_s_att.bias": (1,),
        "fc_affine_v_c_att.weight": (C, H), "fc_affine_v_c_att.bias": (C,),
        "my_tokens": (K, C), "gate_av": (1,),
        "down_sampler.weight": (Cd, C // g, 1, 1), "up_sampler.weight": (C, Cd // g, 1, 1),
    }
    if cfg.use_gate:
        s["gate"] = (1,)
    if cfg.use_bn:
        for n, d in (("bn1", Cd), ("bn2", C)):
            s[f"{n}.weight"] = (d,); s[f"{n}.bias"] = (d,)
            s[f"{n}.running_mean"] = (d,); s[f"{n}.running_var"] = (d,)
    if cfg.ln_before:
        s["ln_before.weight"] = (C,); s["ln_before.bias"] = (C,)
    if cfg.ln_post:
        s["ln_post.weight"] = (C,); s["ln_post.bias"] = (C,)
    if cfg.temporal:
        s["temporal_gated.0.weight"] = (1, C); s["temporal_gated.0.bias"] = (1,)
    return s


def random_params(cfg: AdapterConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Reference-style initialisation with the zero-init gates / BN buffers randomised
    (SURVEY.md §7 hard part 6: a freshly constructed adapter outputs exactly 0)."""
    gen = torch.Generator().manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}
    for name, shp in param_shapes(cfg).items():
        if name in ("gate", "gate_av"):
            t = torch.rand(shp, generator=gen, dtype=torch.float64) * 0.75 + 0.25
        elif name == "my_tokens":
            t = torch.rand(shp, generator=gen, dtype=torch.float64)
        elif name.endswith("running_var"):
            t = torch.rand(shp, generator=gen, dtype=torch.float64) + 0.5
        elif name.endswith("running_mean"):
            t = torch.randn(shp, generator=gen, dtype=torch.float64) * 0.1
        elif name.startswith(("bn", "ln")) and name.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=gen, dtype=torch.float64)
        elif name.startswith(("bn", "ln")) and name.endswith("bias"):
            t = 0.1 * torch.randn(shp, generator=gen, dtype=torch.float64)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            if name.endswith(".bias"):
                wshape = param_shapes(cfg)[name[:-4] + "weight"]
                fan_in = 1
                for d in wshape[1:]:
                    fan_in *= d
            bound = 1.0 / math.sqrt(max(fan_in, 1))
            t = (torch.rand(shp, generator=gen, dtype=torch.float64) * 2 - 1) * bound
        p[name] = t.to(dtype)
    return p


def random_inputs(cfg: AdapterConfig, BT: int, seed: int = 0, dtype=torch.float32):
    gen = torch.Generator().manual_seed(seed + 1000003)
    x = torch.randn(BT, cfg.N, cfg.C, generator=gen, dtype=torch.float64).to(dtype)
    other = torch.randn(BT, cfg.No, cfg.Co, generator=gen, dtype=torch.float64).to(dtype)
    return x, other


def _grouped_pointwise(y: torch.Tensor, w: torch.Tensor, g: int) -> torch.Tensor:
    """1x1 grouped convolution on token-major data: y (BT,N,Cin), w (Cout, Cin/g, 1, 1)."""
    BT, N, Cin = y.shape
    Cout = w.shape[0]
    wg = w.reshape(g, Cout // g, Cin // g)
    yg = y.reshape(BT, N, g, Cin // g)
    return torch.einsum("bngj,goj->bngo", yg, wg).reshape(BT, N, Cout)


def _batchnorm(z, p, name, cfg, training, new_stats):
    """BatchNorm2d over (BT, N) per channel on token-major data (net_trans.py:636,643)."""
    w, b = p[f"{name}.weight"], p[f"{name}.bias"]
    if training:
        mean = z.mean(dim=(0, 1))
        var = z.var(dim=(0, 1), unbiased=False)
        n = z.shape[0] * z.shape[1]
        with torch.no_grad():
            m = cfg.bn_momentum
            unb = var * (n / max(n - 1, 1))
            new_stats[f"{name}.running_mean"] = (1 - m) * p[f"{name}.running_mean"] + m * mean
            new_stats[f"{name}.running_var"] = (1 - m) * p[f"{name}.running_var"] + m * unb
    else:
        mean, var = p[f"{name}.running_mean"], p[f"{name}.running_var"]
    return (z - mean) * torch.rsqrt(var + cfg.eps_bn) * w + b


def resample(cfg: AdapterConfig, p, other: torch.Tensor) -> torch.Tensor:
    """other (BT,No,Co) -> V (BT,N,C)."""
    if cfg.resampler == "conv":
        wc = p["conv_adapter.weight"].reshape(cfg.N, cfg.No)
        u = torch.einsum("nm,bmc->bnc", wc, other) + p["conv_adapter.bias"][None, :, None]
        return u @ p["fc.weight"].t() + p["fc.bias"]
    if cfg.resampler == "bicubic":
        v0 = other @ p["fc.weight"].t() + p["fc.bias"]                   # (BT,No,C)
        hw, s = int(math.sqrt(cfg.No)), int(math.sqrt(cfg.N))
        img = v0.reshape(-1, hw, hw, cfg.C).permute(0, 3, 1, 2)          # 'BF w h c -> BF c w h'
        img = F.interpolate(img, mode="bicubic", size=[s, s])
        return img.reshape(img.shape[0], cfg.C, -1).permute(0, 2, 1)
    raise ValueError(cfg.resampler)


def adapter_forward(cfg: AdapterConfig, p: Dict[str, torch.Tensor], x: torch.Tensor, other: torch.Tensor,
                    training: bool = False, want: bool = False):
    """Returns (out (BT,N,C), spatial_att_maps (BT,1,N), temporal_att_maps (B,T,1,1) | None,
    new_running_stats dict, intermediates dict if want)."""
    BT = x.shape[0]
    inter: Dict[str, torch.Tensor] = {}
    new_stats: Dict[str, torch.Tensor] = {}
    V = resample(cfg, p, other)                                            # (BT,N,C)
    tok = p["my_tokens"]                                                   # (K,C)
    # latent-token attention, no 1/sqrt(d) scaling
    P = torch.softmax(torch.einsum("kc,bnc->bkn", tok, V), dim=-1)         # over n
    Tp = tok[None] + torch.einsum("bkn,bnc->bkc", P, V)
    Q = torch.softmax(torch.einsum("bnc,bkc->bnk", x, Tp), dim=-1)         # over k
    X1 = x + p["gate_av"] * torch.einsum("bnk,bkc->bnc", Q, Tp)
    # channel gate
    obar = V.mean(dim=1)                                                   # (BT,C)
    aq1 = torch.relu(obar @ p["fc_affine_audio_1.weight"].t() + p["fc_affine_audio_1.bias"])
    vq1 = torch.relu(X1 @ p["fc_affine_video_1.weight"].t() + p["fc_affine_video_1.bias"])
    q = (aq1[:, None, :] * vq1).mean(dim=1)
    h = torch.relu(q @ p["fc_affine_bottleneck.weight"].t() + p["fc_affine_bottleneck.bias"])
    m = torch.sigmoid(h @ p["fc_affine_v_c_att.weight"].t() + p["fc_affine_v_c_att.bias"])   # (BT,C)
    # spatial gate
    Xc = X1 * (1 + m[:, None, :])
    vq2 = torch.relu(Xc @ p["fc_affine_video_2.weight"].t() + p["fc_affine_video_2.bias"])
    aq2 = torch.relu(obar @ p["fc_affine_audio_2.weight"].t() + p["fc_affine_audio_2.bias"])
    s = (vq2 * aq2[:, None, :]) @ p["fc_affine_v_s_att.weight"].reshape(-1) + p["fc_affine_v_s_att.bias"]
    sig = torch.sigmoid(s)                                                 # (BT,N)
    smap = torch.softmax(torch.tanh(s), dim=-1)[:, None, :]                # (BT,1,N)
    # temporal gate
    tmap = None
    tau_bt = None
    if cfg.temporal:
        tau = torch.sigmoid(obar.reshape(BT // cfg.T, cfg.T, cfg.C) @ p["temporal_gated.0.weight"].t()
                            + p["temporal_gated.0.bias"])                  # (B,T,1)
        tmap = tau[..., None]
        tau_bt = tau.reshape(BT)
    scale = cfg.alpha * m[:, None, :] + cfg.beta * sig[:, :, None] + (1 - cfg.alpha)
    if cfg.temporal and cfg.gamma != 0.0:
        scale = scale + cfg.gamma * tau_bt[:, None, None]
    Y = X1 * scale
    if cfg.ln_before:
        Y = F.layer_norm(Y, (cfg.C,), p["ln_before.weight"], p["ln_before.bias"], cfg.eps_ln)
    Z = _grouped_pointwise(Y, p["down_sampler.weight"], cfg.g)
    if cfg.use_bn:
        Z = _batchnorm(Z, p, "bn1", cfg, training, new_stats)
    Zr = torch.relu(Z)
    U = _grouped_pointwise(Zr, p["up_sampler.weight"], cfg.g)
    if cfg.use_bn:
        U = _batchnorm(U, p, "bn2", cfg, training, new_stats)
    out = U
    if cfg.gate_before_ln:
        if cfg.use_gate:
            out = p["gate"] * out
        if cfg.ln_post:
            out = F.layer_norm(out, (cfg.C,), p["ln_post.weight"], p["ln_post.bias"], cfg.eps_ln)
    else:
        if cfg.ln_post:
            out = F.layer_norm(out, (cfg.C,), p["ln_post.weight"], p["ln_post.bias"], cfg.eps_ln)
        if cfg.use_gate:
            out = p["gate"] * out
    if want:
        inter.update(V=V, P=P, Tp=Tp, Q=Q, X1=X1, obar=obar, aq1=aq1, q=q, m=m, aq2=aq2, s=s, sig=sig,
                     Y=Y, Z=Z, U=U)
        return out, smap, tmap, new_stats, inter
    return out, smap, tmap, new_stats


def bicubic_matrix(n_in_side: int, n_out_side: int, dtype=torch.float64) -> torch.Tensor:
    """Dense (N_out, N_in) matrix of torch's bicubic F.interpolate (align_corners=False, A=-0.75)
    on a square token grid, obtained by pushing the identity through the reference op.  Used by
    tests to check the engine's own host-side construction."""
    n_in = n_in_side * n_in_side
    eye = torch.eye(n_in, dtype=dtype).reshape(n_in, 1, n_in_side, n_in_side)
    out = F.interpolate(eye, mode="bicubic", size=[n_out_side, n_out_side])
    return out.reshape(n_in, -1).t().contiguous()


# ---- workload tables -----------------------------------------------------------
def ave_l192_layers():
    """The 48 adapter calls of one AVE forward at the reference geometry (Swin-V2-L/192 + HTS-AT/256,
    net_trans.py:781-845, 882-906): list of (name, C, N, Co, No)."""
    vis = [(192, 2304), (384, 576), (768, 144), (1536, 36)]
    aud = [(96, 4096), (192, 1024), (384, 256), (768, 64)]
    pairs = [0, 0, 1, 1, 2, 2, 2, 2, 2, 2, 3, 3]
    out = []
    for li, s in enumerate(pairs):
        for pos in ("p1", "p2"):
            out.append((f"aud_{pos}_{li}", aud[s][0], aud[s][1], vis[s][0], vis[s][1]))
            out.append((f"vis_{pos}_{li}", vis[s][0], vis[s][1], aud[s][0], aud[s][1]))
    return out


def algorithmic_bytes(cfg: AdapterConfig, BT: int, s_act: int = 2, s_w: int = 2) -> int:
    """read x, other and every parameter once; write out and the spatial map."""
    C, N, Co, No, K, r, g = cfg.C, cfg.N, cfg.Co, cfg.No, cfg.K, cfg.r, cfg.g
    P = (N * No + N + Co * C + C + 2 * (C * C + C) + 3 * (C * C // 2 + C // 2) + C // 2 + 1 + C * C // 2 + C
         + K * C + 3 + 2 * C * (C // r) // g + 2 * (C // r) + 6 * C)
    return s_act * BT * (2 * N * C + No * Co + N) + s_w * P


def algorithmic_flops(cfg: AdapterConfig, BT: int) -> int:
    """SURVEY.md §8(d), reference op order."""
    C, N, Co, No, K, r, g = cfg.C, cfg.N, cfg.Co, cfg.No, cfg.K, cfg.r, cfg.g
    return (2 * BT * (N * No * Co + N * Co * C) + 8 * BT * K * N * C + 2 * BT * N * C * C
            + BT * N * C * C + 4 * BT * N * C * (C // r) // g)
