"""bf16 gradient error table: engine (tensor-core path) and torch's own bf16 autocast, both against the fp32 oracle.
Run on the GPU box: python scripts/grad_table.py [variant]"""
import dataclasses
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dgsct_b200.adapter import adapter_backward_raw, adapter_forward_raw   # noqa: E402
from oracle import dgsct_oracle as O                                        # noqa: E402

SHAPES = {"vis_s1": (384, 576, 192, 1024), "vis_s2": (768, 144, 384, 256), "aud_s3": (768, 64, 1536, 36), "aud_s1": (192, 1024, 384, 576),
          "vis_s3": (1536, 36, 768, 64), "aud_s2": (384, 256, 768, 144)}
VARIANTS = {"ave": {}, "s4": dict(use_bn=False, use_gate=False), "ms3": {}, "avqa": dict(K=2, g=4, use_bn=False, use_gate=False)}


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def oracle_grads(cfg, params, x, other, R_out, R_smap, autocast):
    p = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in params.items()}
    xs, os_ = x.clone().requires_grad_(True), other.clone().requires_grad_(True)
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        out, smap, tmap, _ = O.adapter_forward(cfg, p, xs, os_, training=True)
    ((out.float() * R_out).sum() + (smap.float().reshape(R_smap.shape) * R_smap).sum()).backward()
    g = {k: v.grad for k, v in p.items() if v.grad is not None}
    g["x"], g["other"] = xs.grad, os_.grad
    return g, out.detach().float(), smap.detach().float()


def main():
    variant = sys.argv[1] if len(sys.argv) > 1 else "ave"
    res = {}
    for shape, (C, N, Co, No) in SHAPES.items():
        cfg = O.make_config(variant, C, N, Co, No, **VARIANTS[variant])
        params = O.random_params(cfg, seed=21)
        x, other = O.random_inputs(cfg, 10, seed=22)
        g = torch.Generator().manual_seed(23)
        R_out = torch.randn(10, N, C, generator=g).bfloat16().float()
        R_smap = torch.randn(10, N, generator=g).bfloat16().float()
        xq, oq = x.bfloat16(), other.bfloat16()
        p = {k: v.float().cuda().contiguous() for k, v in params.items()}
        cd = dataclasses.asdict(cfg)
        out, smap, _ = adapter_forward_raw(cd, {k: v.clone() for k, v in p.items()}, xq.cuda(), oq.cuda(), training=True)
        dx, do, grads = adapter_backward_raw(cd, p, xq.cuda(), oq.cuda(), R_out.bfloat16().cuda(), R_smap.bfloat16().cuda(), None, training=True)
        torch.cuda.synchronize()
        grads = {k: v.cpu() for k, v in grads.items()}
        grads["x"], grads["other"] = dx.float().cpu(), do.float().cpu()
        ref, ro, rs = oracle_grads(cfg, params, xq.float(), oq.float(), R_out, R_smap, False)
        ac, _, _ = oracle_grads(cfg, params, xq.float(), oq.float(), R_out, R_smap, True)
        row = {"fwd_out": (rel(out.float().cpu(), ro), None), "fwd_smap": (rel(smap.float().cpu(), rs.reshape(smap.shape)), None)}
        for k, v in ref.items():
            if k in grads and float(v.abs().max()) > 1e-6:
                row[k] = (rel(grads[k].reshape(v.shape), v), rel(ac[k].float(), v) if k in ac else None)
        res[shape] = row
        print(f"== {variant} {shape}  (engine bf16 | torch bf16 autocast)  rel-L2 vs fp32 oracle")
        for k, (e, a) in sorted(row.items(), key=lambda kv: -kv[1][0]):
            print(f"   {k:34s} {e:9.2e}   {a if a is None else format(a, '9.2e')}")
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        json.dump(res, open(os.path.join(out_dir, f"grad_table_{variant}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
