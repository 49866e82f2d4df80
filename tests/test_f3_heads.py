"""token pooling with the adapters' spatial maps and the TemporalAttention gate tail.
CPU: the oracle restatement against golden vectors produced by the reference's own TemporalAttention class.
GPU: the engine kernels (through the C ABI) against the oracle and the golden vectors."""
import os

import numpy as np
import pytest
import torch

from oracle import heads_oracle as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "f3_temporal_gates.npz")


def _gold():
    z = np.load(GOLD)
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def test_oracle_gates_match_reference_temporal_attention():
    g = _gold()
    leaves = {k: g[k].clone().requires_grad_(True) for k in ("a_kv", "v_kv", "vq", "aq", "wa", "ba", "wv", "bv")}
    vq_o, aq_o, av = H.temporal_gates(leaves["a_kv"], leaves["v_kv"], leaves["vq"], leaves["aq"], leaves["wa"], leaves["ba"],
                                      leaves["wv"], leaves["bv"], float(g["gamma"]))
    assert _rel(vq_o, g["vq_out"]) < 1e-12 and _rel(aq_o, g["aq_out"]) < 1e-12 and _rel(av, g["av"]) < 1e-12
    ((vq_o * g["R_vq"]).sum() + (aq_o * g["R_aq"]).sum() + (av * g["R_av"]).sum()).backward()
    # (the encoder outputs a_kv / v_kv also feed the decoders in the reference, so only tensors that feed the tail alone
    #  have comparable golden gradients)
    for k in ("vq", "aq", "wa", "ba", "wv", "bv"):
        assert _rel(leaves[k].grad, g["grad_" + k]) < 1e-10, k


def test_oracle_pooling_is_bmm():
    torch.manual_seed(0)
    m = torch.softmax(torch.randn(6, 1, 36, dtype=torch.float64), -1)
    f = torch.randn(6, 36, 48, dtype=torch.float64)
    assert torch.allclose(H.pool_tokens(m, f), torch.bmm(m, f), atol=1e-14)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_engine_gates_match_oracle_and_golden(dtype):
    from dgsct_b200.heads import TemporalGates
    g = _gold()
    mod = TemporalGates(256, gamma=float(g["gamma"])).cuda()
    with torch.no_grad():
        mod.audio_gated[0].weight.copy_(g["wa"]); mod.audio_gated[0].bias.copy_(g["ba"])
        mod.video_gated[0].weight.copy_(g["wv"]); mod.video_gated[0].bias.copy_(g["bv"])
    ins = {k: g[k].to(dtype).cuda().requires_grad_(True) for k in ("a_kv", "v_kv", "vq", "aq")}
    vq_o, aq_o, av = mod(ins["a_kv"], ins["v_kv"], ins["vq"], ins["aq"])
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert vq_o.shape == g["vq_out"].shape and av.shape == g["av"].shape
    assert _rel(vq_o.float().cpu(), g["vq_out"]) < tol and _rel(aq_o.float().cpu(), g["aq_out"]) < tol and _rel(av.float().cpu(), g["av"]) < tol
    R = [g[k].to(dtype).cuda() for k in ("R_vq", "R_aq", "R_av")]
    torch.autograd.backward([vq_o, aq_o, av], R)
    torch.cuda.synchronize()
    # oracle gradients on the same (rounded) inputs
    leaves = {k: ins[k].detach().float().cpu().double().requires_grad_(True) for k in ins}
    ws = {k: g[k].clone().requires_grad_(True) for k in ("wa", "ba", "wv", "bv")}
    o = H.temporal_gates(leaves["a_kv"], leaves["v_kv"], leaves["vq"], leaves["aq"], ws["wa"], ws["ba"], ws["wv"], ws["bv"], float(g["gamma"]))
    sum((a * r.float().cpu().double()).sum() for a, r in zip(o, R)).backward()
    gtol = 1e-4 if dtype == torch.float32 else 2e-2
    for k in ins:
        assert _rel(ins[k].grad.float().cpu(), leaves[k].grad) < gtol, k
    got = {"wa": mod.audio_gated[0].weight.grad, "ba": mod.audio_gated[0].bias.grad, "wv": mod.video_gated[0].weight.grad,
           "bv": mod.video_gated[0].bias.grad}
    for k, v in got.items():
        assert v.shape == ws[k].shape and _rel(v.float().cpu(), ws[k].grad) < gtol, k
    if dtype == torch.float32:
        for k in ("vq", "aq"):
            assert _rel(ins[k].grad.cpu(), g["grad_" + k]) < 1e-5, k          # straight against the reference class


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(80, 36, 1536), (80, 64, 768), (20, 2304, 192)])
def test_engine_pooling_matches_bmm(dtype, shape):
    from dgsct_b200.heads import pool_tokens
    BT, N, C = shape
    gen = torch.Generator().manual_seed(3)
    m = torch.softmax(torch.randn(BT, 1, N, generator=gen), -1).to(dtype).cuda().requires_grad_(True)
    f = torch.randn(BT, N, C, generator=gen).to(dtype).cuda().requires_grad_(True)
    out = pool_tokens(m, f)
    assert out.shape == (BT, 1, C)
    R = torch.randn(BT, 1, C, generator=gen).to(dtype).cuda()
    out.backward(R)
    torch.cuda.synchronize()
    m2, f2 = m.detach().float().cpu().double().requires_grad_(True), f.detach().float().cpu().double().requires_grad_(True)
    ref = H.pool_tokens(m2, f2)
    (ref * R.float().cpu().double()).sum().backward()
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert _rel(out.float().cpu(), ref) < tol
    assert _rel(m.grad.float().cpu(), m2.grad) < tol and _rel(f.grad.float().cpu(), f2.grad) < tol


@pytest.mark.gpu
def test_pooling_consumes_the_adapters_second_output():
    """the chain of net_trans.py:901,922: last visual adapter -> spatial_att_maps -> pooled visual feature"""
    import types
    import dgsct_b200
    from dgsct_b200.heads import pool_tokens
    from oracle import dgsct_oracle as O
    C, N, Co, No = 1536, 36, 768, 64
    opt = types.SimpleNamespace(is_multimodal=1, num_conv_group=2, is_before_layernorm=1, is_post_layernorm=1)
    ad = dgsct_b200.VisualAdapter(C, C, "bottleneck", reduction_factor=8, opt=opt, num_tk=32, conv_dim_in=No, conv_dim_out=N,
                                  linear_in=Co, linear_out=C).cuda().eval()
    cfg = O.make_config("ave", C, N, Co, No)
    params = O.random_params(cfg, seed=9)
    ad.load_state_dict(params, strict=False)
    f_v, f_a = torch.randn(10, N, C, device="cuda"), torch.randn(10, No, Co, device="cuda")
    with torch.no_grad():
        res, smap = ad(f_v.permute(0, 2, 1).unsqueeze(-1), f_a.permute(0, 2, 1).unsqueeze(-1))
        f_v2 = f_v + res.squeeze(-1).permute(0, 2, 1)
        pooled = pool_tokens(smap, f_v2)
    ro, rs, _, _ = O.adapter_forward(cfg, params, f_v.cpu(), f_a.cpu())
    ref = torch.bmm(rs, f_v.cpu() + ro)
    assert _rel(pooled.cpu(), ref) < 1e-4
