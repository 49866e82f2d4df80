"""Parity of the CUDA engine (through the C ABI) against the golden reference fixtures and the oracle.
Tolerances are the north-star's: 1e-3 relative for fp32 I/O, 1e-2 relative (rel-L2) for bf16 I/O."""
import dataclasses
import types

import pytest
import torch

import dgsct_b200
from dgsct_b200.adapter import adapter_forward_raw
from oracle import dgsct_oracle as O
from tests.golden_util import case_names, load_case, rel_l2

pytestmark = pytest.mark.gpu
TOL_F32, TOL_BF16 = 1e-3, 1e-2


def cfg_dict(cfg):
    return dataclasses.asdict(cfg)


def to_dev(params, dtype=torch.float32):
    return {k: v.to(dtype).cuda().contiguous() for k, v in params.items()}


@pytest.mark.parametrize("name", case_names())
@pytest.mark.parametrize("training", [False, True])
def test_fp32_matches_reference_golden(name, training):
    cfg, meta, params, t = load_case(name)
    p = to_dev(params)
    x, other = t["x"].float().cuda(), t["other"].float().cuda()
    out, smap, tmap = adapter_forward_raw(cfg_dict(cfg), p, x, other, training=training)
    torch.cuda.synchronize()
    tag = "train" if training else "eval"
    assert rel_l2(out.cpu(), t[f"{tag}/out"]) < TOL_F32
    assert rel_l2(smap.cpu(), t[f"{tag}/smap"].reshape(smap.shape)) < TOL_F32
    # much tighter in practice: the fp32 path is exact fp32 arithmetic
    assert rel_l2(out.cpu(), t[f"{tag}/out"]) < 2e-5
    if cfg.temporal:
        assert rel_l2(tmap.cpu(), t[f"{tag}/tmap"].reshape(-1)) < 1e-5
    if training and cfg.use_bn:
        for bn in ("bn1", "bn2"):
            for s in ("running_mean", "running_var"):
                assert rel_l2(p[f"{bn}.{s}"].cpu(), t[f"train/{bn}.{s}"]) < 1e-5, (bn, s)


@pytest.mark.parametrize("name", ["ave_vis", "s4_down", "fewshot_aud"])
def test_taps_match_oracle_intermediates(name):
    cfg, meta, params, t = load_case(name)
    p = to_dev(params)
    x, other = t["x"].float().cuda(), t["other"].float().cuda()
    out, smap, tmap, taps = adapter_forward_raw(cfg_dict(cfg), p, x, other, training=False, want_taps=True)
    torch.cuda.synchronize()
    _, _, _, _, inter = O.adapter_forward(cfg, params, t["x"], t["other"], training=False, want=True)
    for k in ("V", "Tp", "X1", "obar", "m", "s", "Y"):
        assert rel_l2(taps[k].cpu(), inter[k]) < 2e-5, k


SHAPES = {  # SURVEY.md §8(d) shape classes (Swin-V2-L/192 + HTS-AT/256), plus B/224 and CLIP few-shot ones
    "vis_s0": (192, 2304, 96, 4096), "vis_s1": (384, 576, 192, 1024), "vis_s2": (768, 144, 384, 256),
    "vis_s3": (1536, 36, 768, 64), "aud_s0": (96, 4096, 192, 2304), "aud_s1": (192, 1024, 384, 576),
    "aud_s2": (384, 256, 768, 144), "aud_s3": (768, 64, 1536, 36),
    "b224_s1": (256, 784, 192, 1024), "b224_s3": (1024, 49, 768, 64), "clip": (768, 50, 96, 4096),
}


def _run_shape(shape, dtype, BT, training, variant="ave", **kw):
    C, N, Co, No = SHAPES[shape]
    cfg = O.make_config(variant, C, N, Co, No, **kw)
    params = O.random_params(cfg, seed=7)
    x, other = O.random_inputs(cfg, BT, seed=11)
    xq, oq = x.to(dtype), other.to(dtype)
    p = to_dev(params)
    out, smap, tmap = adapter_forward_raw(cfg_dict(cfg), p, xq.cuda(), oq.cuda(), training=training)
    torch.cuda.synchronize()
    ro, rs, rt, rstats = O.adapter_forward(cfg, params, xq.float(), oq.float(), training=training)
    return cfg, out.float().cpu(), smap.float().cpu(), ro, rs.reshape(rs.shape[0], -1), p, rstats


@pytest.mark.parametrize("shape", list(SHAPES))
def test_real_shapes_fp32(shape):
    BT = 4 if SHAPES[shape][1] >= 2304 else 10
    cfg, out, smap, ro, rs, *_ = _run_shape(shape, torch.float32, BT, False)
    assert rel_l2(out, ro) < TOL_F32 and rel_l2(smap, rs) < TOL_F32
    assert rel_l2(out, ro) < 5e-5
    assert float((smap.sum(-1) - 1).abs().max()) < 1e-4


@pytest.mark.parametrize("shape", list(SHAPES))
def test_real_shapes_bf16(shape):
    BT = 4 if SHAPES[shape][1] >= 2304 else 10
    cfg, out, smap, ro, rs, *_ = _run_shape(shape, torch.bfloat16, BT, False)
    assert rel_l2(out, ro) < TOL_BF16, rel_l2(out, ro)
    assert rel_l2(smap, rs) < TOL_BF16, rel_l2(smap, rs)


@pytest.mark.parametrize("shape", ["vis_s0", "vis_s2", "aud_s1", "aud_s3", "clip"])
def test_bf16_tensor_core_taps(shape):
    """Stage-by-stage parity of the tcgen05 pipeline (bf16 I/O) against oracle intermediates."""
    C, N, Co, No = SHAPES[shape]
    BT = 4 if N >= 2304 else 10
    cfg = O.make_config("ave", C, N, Co, No)
    params = O.random_params(cfg, seed=5)
    x, other = O.random_inputs(cfg, BT, seed=6)
    xq, oq = x.bfloat16(), other.bfloat16()
    p = to_dev(params)
    out, smap, tmap, taps = adapter_forward_raw(cfg_dict(cfg), p, xq.cuda(), oq.cuda(), want_taps=True)
    torch.cuda.synchronize()
    from dgsct_b200 import _lib as L
    assert L.lib().dgsct_tc_error_flag(1) == 0
    ro, rs, _, _, inter = O.adapter_forward(cfg, params, xq.float(), oq.float(), want=True)
    errs = {k: rel_l2(taps[k].cpu(), inter[k]) for k in ("V", "Tp", "X1", "obar", "m", "s", "Y")}
    errs["out"] = rel_l2(out.float().cpu(), ro)
    errs["smap"] = rel_l2(smap.float().cpu(), rs.reshape(smap.shape))
    print(shape, {k: f"{v:.2e}" for k, v in errs.items()})
    for k, v in errs.items():
        assert v < TOL_BF16, (k, errs)


@pytest.mark.parametrize("shape", ["vis_s1", "aud_s2"])
def test_bf16_tensor_core_vs_simt_twin(shape):
    C, N, Co, No = SHAPES[shape]
    cfg = O.make_config("ave", C, N, Co, No)
    p = to_dev(O.random_params(cfg, seed=5))
    x, other = O.random_inputs(cfg, 10, seed=6)
    xq, oq = x.bfloat16().cuda(), other.bfloat16().cuda()
    o1, s1, _ = adapter_forward_raw(cfg_dict(cfg), p, xq, oq)
    o2, s2, _ = adapter_forward_raw(cfg_dict(cfg), p, xq, oq, force_simt=True)
    packed = adapter_forward_raw(cfg_dict(cfg), p, xq, oq, pack_only=True)
    o3, s3, _ = adapter_forward_raw(cfg_dict(cfg), p, xq, oq, packed=packed)
    torch.cuda.synchronize()
    assert rel_l2(o1.float().cpu(), o2.float().cpu()) < TOL_BF16
    assert rel_l2(o1.float().cpu(), o3.float().cpu()) < 2e-3     # pre-packed weights: same kernels


@pytest.mark.parametrize("shape,dtype", [("vis_s2", torch.float32), ("aud_s2", torch.bfloat16), ("vis_s3", torch.bfloat16),
                                         ("aud_s0", torch.bfloat16)])
def test_train_mode_batch_stats(shape, dtype):
    cfg, out, smap, ro, rs, p, rstats = _run_shape(shape, dtype, 10, True)
    tol = TOL_F32 if dtype == torch.float32 else TOL_BF16
    assert rel_l2(out, ro) < tol
    for k, v in rstats.items():
        assert rel_l2(p[k].cpu(), v) < tol, k


VARIANT_CASES = [("s4", dict(use_bn=False, use_gate=False)), ("ms3", {}), ("pretrain", dict(K=2)),
                 ("avqa", dict(K=2, g=4, use_bn=False, use_gate=False)), ("fewshot", dict(K=2, alpha=0.7, beta=0.2, gamma=0.4))]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_variants_at_real_shape(dtype):
    for variant, kw in VARIANT_CASES:
        shape = "vis_s2"
        cfg, out, smap, ro, rs, *_ = _run_shape(shape, dtype, 10, False, variant=variant, **kw)
        # bf16: every variant meets the north-star bar (1e-2 rel-L2).  Round 1 needed 2e-2 for S4: its bicubic interpolation ran
        # as a dense matrix whose COEFFICIENTS were rounded to bf16 (9.5e-3 of the 1.1e-2); it is now a 16-tap stencil with
        # fp32 coefficients (rowwise.cuh spmm_frames_kernel).
        tol = 5e-5 if dtype == torch.float32 else TOL_BF16
        assert rel_l2(out, ro) < tol, (variant, rel_l2(out, ro))
        assert rel_l2(smap, rs) < tol, (variant, rel_l2(smap, rs))


def test_module_contract_and_strided_inputs():
    opt = types.SimpleNamespace(is_multimodal=1, num_conv_group=2, is_before_layernorm=1, is_post_layernorm=1)
    C, N, Co, No = SHAPES["vis_s3"]
    ad = dgsct_b200.VisualAdapter(C, C, "bottleneck", reduction_factor=8, opt=opt, use_bn=True, use_gate=True, num_tk=32,
                                  conv_dim_in=No, conv_dim_out=N, linear_in=Co, linear_out=C).cuda().eval()
    cfg = O.make_config("ave", C, N, Co, No)
    params = O.random_params(cfg, seed=3)
    ad.load_state_dict({k: v for k, v in params.items()}, strict=False)
    f_v = torch.randn(10, N, C, device="cuda")
    f_a = torch.randn(10, No, Co, device="cuda")
    with torch.no_grad():
        out, smap = ad(f_v.permute(0, 2, 1).unsqueeze(-1), f_a.permute(0, 2, 1).unsqueeze(-1))   # net_trans.py:891
        # genuinely channel-major (contiguous (BT,C,N,1)) inputs must give the same answer
        out2, smap2 = ad(f_v.permute(0, 2, 1).contiguous().unsqueeze(-1), f_a.permute(0, 2, 1).contiguous().unsqueeze(-1))
    assert out.shape == (10, C, N, 1) and smap.shape == (10, 1, N)
    # (not bit-equal: the per-frame reductions use fp32 atomics, so summation order varies run to run)
    assert rel_l2(out, out2) < 1e-5 and rel_l2(smap, smap2) < 1e-5
    ro, rs, _, _ = O.adapter_forward(cfg, params, f_v.cpu(), f_a.cpu())
    assert rel_l2(out.squeeze(-1).permute(0, 2, 1).cpu(), ro) < 5e-5
    # a fresh adapter (zero gates) outputs exactly zero, as in the reference (SURVEY.md §7 hard part 6)
    ad0 = dgsct_b200.VisualAdapter(C, C, "bottleneck", reduction_factor=8, opt=opt, num_tk=32, conv_dim_in=No,
                                   conv_dim_out=N, linear_in=Co, linear_out=C).cuda().eval()
    with torch.no_grad():
        z, _ = ad0(f_v.permute(0, 2, 1).unsqueeze(-1), f_a.permute(0, 2, 1).unsqueeze(-1))
    assert float(z.abs().max()) == 0.0
    # train mode bumps num_batches_tracked like nn.BatchNorm2d
    ad.train()
    with torch.no_grad():
        ad(f_v.permute(0, 2, 1).unsqueeze(-1), f_a.permute(0, 2, 1).unsqueeze(-1))
    assert int(ad.bn1.num_batches_tracked) == 1
