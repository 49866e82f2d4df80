"""Pin oracle/dgsct_oracle.py against the fixtures generated from the reference's own classes
(tests/golden/make_golden.py).  CPU only."""
import pytest
import torch

from oracle import dgsct_oracle as O
from tests.golden_util import case_names, load_case, rel_l2


def _run(cfg, params, t, training, dtype):
    p = {k: v.to(dtype).clone().requires_grad_(v.is_floating_point()) for k, v in params.items()}
    x = t["x"].to(dtype).clone().requires_grad_(True)
    other = t["other"].to(dtype).clone().requires_grad_(True)
    out, smap, tmap, new_stats = O.adapter_forward(cfg, p, x, other, training=training)
    loss = (out * t["R_out"].to(dtype)).sum() + (smap * t["R_smap"].to(dtype)).sum()
    if tmap is not None and cfg.gamma != 0.0:
        loss = loss + (tmap * t["R_tmap"].to(dtype)).sum()
    loss.backward()
    return out, smap, tmap, new_stats, x.grad, other.grad, p


@pytest.mark.parametrize("name", case_names())
@pytest.mark.parametrize("training", [False, True])
def test_oracle_matches_reference_fp64(name, training):
    cfg, meta, params, t = load_case(name)
    tag = "train" if training else "eval"
    out, smap, tmap, new_stats, gx, go, p = _run(cfg, params, t, training, torch.float64)
    assert rel_l2(out, t[f"{tag}/out"]) < 1e-11
    assert rel_l2(smap, t[f"{tag}/smap"]) < 1e-11
    assert tuple(smap.shape) == tuple(t[f"{tag}/smap"].shape)
    assert torch.allclose(smap.sum(-1), torch.ones_like(smap.sum(-1)))
    if f"{tag}/tmap" in t and cfg.temporal:
        assert rel_l2(tmap, t[f"{tag}/tmap"]) < 1e-11
    assert rel_l2(gx, t[f"{tag}/grad_x"]) < 1e-9
    assert rel_l2(go, t[f"{tag}/grad_other"]) < 1e-9
    n_checked = 0
    for k, v in t.items():
        if k.startswith(f"{tag}/grad/"):
            pn = k[len(f"{tag}/grad/"):]
            if pn not in p:      # parameters the restatement does not own (dead: gate_tk, fc_caption, ...)
                assert float(v.abs().max()) == 0.0 or pn.startswith(("temporal_gated", "conv_adapter")), pn
                continue
            g = p[pn].grad
            if g is None:
                assert float(v.abs().max()) == 0.0, pn
                continue
            assert rel_l2(g, v) < 1e-8 or float((g - v).abs().max()) < 1e-12, (pn, rel_l2(g, v))
            n_checked += 1
    assert n_checked >= 18
    if training and cfg.use_bn:
        for bn in ("bn1", "bn2"):
            for s in ("running_mean", "running_var"):
                assert rel_l2(new_stats[f"{bn}.{s}"], t[f"train/{bn}.{s}"]) < 1e-11


@pytest.mark.parametrize("name", case_names())
def test_oracle_fp32_floor(name):
    cfg, meta, params, t = load_case(name)
    out, smap, *_ = _run(cfg, params, t, False, torch.float32)
    assert rel_l2(out, t["eval/out"]) < 2e-5
    assert rel_l2(smap, t["eval/smap"]) < 2e-5


def test_param_shapes_match_reference_state_dict():
    for name in case_names():
        cfg, meta, params, t = load_case(name)
        shapes = O.param_shapes(cfg)
        assert set(shapes) == set(params)
        for k, v in params.items():
            assert tuple(v.shape) == shapes[k]
        # keys the reference has beyond ours are the documented dead ones
        for k in meta["extra_state_keys"]:
            assert k.startswith(("gate_tk", "fc_caption", "temporal_gated", "bn1.num_batches", "bn2.num_batches")), k


def test_bicubic_matrix_is_the_interpolation():
    import math
    cfg, meta, params, t = load_case("s4_down")
    M = O.bicubic_matrix(int(math.sqrt(cfg.No)), int(math.sqrt(cfg.N)))
    v0 = t["other"] @ params["fc.weight"].t() + params["fc.bias"]
    V = O.resample(cfg, params, t["other"])
    assert rel_l2(torch.einsum("nm,bmc->bnc", M, v0), V) < 1e-12


def test_workload_totals_match_survey():
    tot_b = tot_f = 0
    for (_, C, N, Co, No) in O.ave_l192_layers():
        cfg = O.make_config("ave", C, N, Co, No)
        tot_b += O.algorithmic_bytes(cfg, 80)
        tot_f += O.algorithmic_flops(cfg, 80)
    assert len(O.ave_l192_layers()) == 48
    assert abs(tot_b / 1e9 - 4.193) < 0.01
    assert abs(tot_f / 1e12 - 3.432) < 0.01


def test_torch_cpu_bn_eval_backward_bug_is_worked_around():
    """torch 2.11 CPU: eval-mode F.batch_norm backward is wrong for a permuted grad_output (the layout
    ln_post hands to bn2 in the reference).  The oracle's BN is explicit arithmetic, so its gradient is
    the analytic one whatever the layout; make_golden.py forces contiguous grads into the reference's BN
    modules and cross-checks with finite differences."""
    cfg = O.make_config("ave", 16, 25, 32, 16, K=4, r=4)
    p = O.random_params(cfg, 0, torch.float64)
    u = torch.randn(10, 25, 16, dtype=torch.float64, requires_grad=True)
    g = torch.randn(10, 16, 25, dtype=torch.float64).permute(0, 2, 1)      # non-contiguous grad
    y = O._batchnorm(u, p, "bn2", cfg, False, {})
    gi, = torch.autograd.grad(y, u, g)
    sc = p["bn2.weight"] / torch.sqrt(p["bn2.running_var"] + cfg.eps_bn)
    assert rel_l2(gi, g * sc) < 1e-14
