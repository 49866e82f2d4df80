"""Generate golden vectors from the REFERENCE's own VisualAdapter classes.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

For each task tree the class source is located with ``ast`` inside the read-only reference file and
executed as-is (no copy is written anywhere); the instance is loaded with seeded non-trivial
parameters (gates, BN buffers randomised -- a fresh adapter outputs exactly 0), run in fp64 in eval
and train mode, and back-propagated through a fixed random linear functional of its outputs.  Inputs,
parameters, outputs, updated BN buffers and all gradients are stored in ``tests/golden/<case>.npz``.
The fixtures pin ``oracle/dgsct_oracle.py`` (tests/test_oracle_golden.py).
"""
import ast
import json
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from einops import rearrange, repeat

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import dgsct_oracle as O  # noqa: E402

REF = "/root/reference"
SOURCES = {
    "ave": "DG-SCT/AVE/nets/net_trans.py",
    "avqa": "DG-SCT/AVQA/net_grd_avst/net_avst.py",
    "avvp": "DG-SCT/AVVP/nets/mgn.py",
    "s4": "DG-SCT/AVS/avs_scripts/avs_s4/model/PVT_AVSModel.py",
    "ms3": "DG-SCT/AVS/avs_scripts/avs_ms3/model/PVT_AVSModel.py",
    "pretrain": "pretrain/nets/net_trans.py",
    "fewshot": "few-shot/nets/net_trans.py",
}

# (case name, variant, dims, extra config)
CASES = [
    ("ave_vis", "ave", dict(C=32, N=16, Co=16, No=25), dict(K=4, r=8, g=2)),
    ("ave_aud", "ave", dict(C=16, N=25, Co=32, No=16), dict(K=4, r=4, g=2)),
    ("ave_nobn", "ave", dict(C=32, N=9, Co=16, No=16), dict(K=3, r=8, g=2, use_bn=False, use_gate=False)),
    ("avqa_vis", "avqa", dict(C=32, N=16, Co=16, No=25), dict(K=2, r=8, g=4, use_bn=False, use_gate=False)),
    ("avvp_vis", "avvp", dict(C=32, N=16, Co=16, No=25), dict(K=4, r=8, g=2)),
    ("s4_down", "s4", dict(C=32, N=16, Co=16, No=36), dict(K=4, r=8, g=2)),
    ("s4_up", "s4", dict(C=16, N=36, Co=32, No=16), dict(K=4, r=4, g=2, use_bn=False, use_gate=False)),
    ("ms3_vis", "ms3", dict(C=32, N=16, Co=16, No=25), dict(K=4, r=8, g=2)),
    ("pretrain_vis", "pretrain", dict(C=32, N=10, Co=16, No=16), dict(K=2, r=8, g=2)),
    ("fewshot_aud", "fewshot", dict(C=16, N=16, Co=32, No=10), dict(K=2, r=4, g=2, alpha=0.7, beta=0.2, gamma=0.4)),
]
BT = 10


def load_reference_class(variant):
    path = os.path.join(REF, SOURCES[variant])
    src = open(path).read()
    tree = ast.parse(src)
    node = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "VisualAdapter"][0]
    mod = ast.Module(body=[node], type_ignores=[])
    ns = dict(torch=torch, nn=nn, F=F, math=math, rearrange=rearrange, repeat=repeat)
    exec(compile(mod, path, "exec"), ns)
    return ns["VisualAdapter"], f"{SOURCES[variant]}:{node.lineno}-{node.end_lineno}"


def build_reference(variant, cfg):
    cls, where = load_reference_class(variant)
    opt = types.SimpleNamespace(is_multimodal=1, num_conv_group=cfg.g, is_before_layernorm=int(cfg.ln_before),
                                is_post_layernorm=int(cfg.ln_post), num_tokens=cfg.K,
                                alpha=cfg.alpha, beta=cfg.beta, gamma=cfg.gamma)
    kw = dict(input_dim=cfg.C, output_dim=cfg.C, adapter_kind="bottleneck", reduction_factor=cfg.r, opt=opt,
              use_bn=cfg.use_bn, use_gate=cfg.use_gate, conv_dim_in=cfg.No, conv_dim_out=cfg.N,
              linear_in=cfg.Co, linear_out=cfg.C)
    if variant in ("ave", "avvp", "pretrain", "fewshot"):
        kw["num_tk"] = cfg.K
    ref = cls(**kw).double()
    # torch 2.11 CPU bug (found while pinning the oracle, see tests/test_oracle_golden.py::
    # test_torch_cpu_bn_eval_backward_bug_is_worked_around): eval-mode batch_norm backward returns a wrong
    # grad_input when grad_output is a permuted (non-contiguous) tensor, which is what ln_post's permute
    # hands to bn2.  The hook only changes the memory format of the incoming gradient, not its values;
    # finite differences of the reference forward (checked below) confirm the hooked gradients.
    for m in ref.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.register_forward_hook(lambda mod, i, o: o.register_hook(lambda g: g.contiguous()) and None
                                    if o.requires_grad else None)
    return ref, where


def run(ref, variant, cfg, x, other, R, training):
    ref.train(training)
    xs = x.clone().requires_grad_(True)
    os_ = other.clone().requires_grad_(True)
    res = ref(xs.permute(0, 2, 1).unsqueeze(-1), os_.permute(0, 2, 1).unsqueeze(-1))
    out = res[0].squeeze(-1).permute(0, 2, 1)
    smap = res[1]
    tmap = res[2] if len(res) > 2 else None
    loss = (out * R["R_out"]).sum() + (smap * R["R_smap"]).sum()
    if tmap is not None and cfg.gamma != 0.0:
        loss = loss + (tmap * R["R_tmap"]).sum()
    for prm in ref.parameters():
        prm.grad = None
    loss.backward()
    grads = {n: prm.grad.detach().clone() for n, prm in ref.named_parameters() if prm.grad is not None}
    return out.detach(), smap.detach(), None if tmap is None else tmap.detach(), xs.grad, os_.grad, grads


def fd_check(ref, cfg, x, other, R, training, gx, go, n=6, eps=1e-6):
    """Central finite differences of the reference FORWARD at a few coordinates vs its autograd."""
    sd = {k: v.clone() for k, v in ref.state_dict().items()}

    def f(xx, oo):
        ref.load_state_dict(sd)
        ref.train(training)
        with torch.no_grad():
            res = ref(xx.permute(0, 2, 1).unsqueeze(-1), oo.permute(0, 2, 1).unsqueeze(-1))
        val = (res[0].squeeze(-1).permute(0, 2, 1) * R["R_out"]).sum() + (res[1] * R["R_smap"]).sum()
        if len(res) > 2 and cfg.gamma != 0.0:
            val = val + (res[2] * R["R_tmap"]).sum()
        return val.item()

    gen = torch.Generator().manual_seed(7)
    for which, base, grad in (("x", x, gx), ("other", other, go)):
        for _ in range(n):
            idx = tuple(int(torch.randint(0, s, (1,), generator=gen)) for s in base.shape)
            tp, tm = base.clone(), base.clone()
            tp[idx] += eps
            tm[idx] -= eps
            fd = (f(tp, other) - f(tm, other)) / (2 * eps) if which == "x" else (f(x, tp) - f(x, tm)) / (2 * eps)
            assert abs(fd - grad[idx].item()) <= 1e-5 * max(1.0, abs(fd)), (which, idx, fd, grad[idx].item())
    ref.load_state_dict(sd)


def main():
    torch.manual_seed(0)
    index = {}
    for ci, (name, variant, dims, extra) in enumerate(CASES):
        cfg = O.make_config(variant, **dims, **extra)
        # the reference variants hard-code T (10; 5 in AVS) and, except few-shot, alpha/beta/gamma
        ref, where = build_reference(variant, cfg)
        params = O.random_params(cfg, seed=100 + ci, dtype=torch.float64)
        sd = ref.state_dict()
        for k, v in params.items():
            assert k in sd and tuple(sd[k].shape) == tuple(v.shape), (name, k, sd.get(k, torch.empty(0)).shape, v.shape)
        missing = ref.load_state_dict(params, strict=False)
        assert not missing.unexpected_keys, missing
        x, other = O.random_inputs(cfg, BT, seed=200 + ci, dtype=torch.float64)
        gen = torch.Generator().manual_seed(300 + ci)
        R = dict(R_out=torch.randn(BT, cfg.N, cfg.C, generator=gen, dtype=torch.float64),
                 R_smap=torch.randn(BT, 1, cfg.N, generator=gen, dtype=torch.float64),
                 R_tmap=torch.randn(BT // cfg.T, cfg.T, 1, 1, generator=gen, dtype=torch.float64))
        blob = {"x": x, "other": other, **R}
        for k, v in params.items():
            blob["param/" + k] = v
        for training in (False, True):
            ref.load_state_dict(params, strict=False)
            out, smap, tmap, gx, go, grads = run(ref, variant, cfg, x, other, R, training)
            tag = "train" if training else "eval"
            fd_check(ref, cfg, x, other, R, training, gx, go)
            blob[f"{tag}/out"] = out
            blob[f"{tag}/smap"] = smap
            if tmap is not None:
                blob[f"{tag}/tmap"] = tmap
            blob[f"{tag}/grad_x"] = gx
            blob[f"{tag}/grad_other"] = go
            for k, v in grads.items():
                blob[f"{tag}/grad/{k}"] = v
            if training and cfg.use_bn:
                sd = ref.state_dict()
                for bn in ("bn1", "bn2"):
                    blob[f"train/{bn}.running_mean"] = sd[f"{bn}.running_mean"].clone()
                    blob[f"train/{bn}.running_var"] = sd[f"{bn}.running_var"].clone()
        meta = dict(variant=variant, cfg={k: getattr(cfg, k) for k in cfg.__dataclass_fields__}, BT=BT,
                    reference_class=where, torch=torch.__version__,
                    extra_state_keys=sorted(set(ref.state_dict().keys()) - set(params.keys())))
        blob_np = {k: v.numpy() for k, v in blob.items()}
        blob_np["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **blob_np)
        index[name] = meta
        print(name, where, "out rms %.4f" % blob["eval/out"].pow(2).mean().sqrt().item(),
              "params", sum(v.numel() for v in params.values()))
    json.dump(index, open(os.path.join(HERE, "index.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
