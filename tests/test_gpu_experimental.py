"""Default-off experiments (compiled, awaiting their first hardware run): each flag must reproduce the default path.
Marked ``gpu_experimental`` -- NOT part of ``-m gpu``; run with ``pytest -m gpu_experimental`` before flipping a default."""
import dataclasses
import os

import pytest
import torch

from dgsct_b200.adapter import adapter_backward_raw, adapter_forward_raw
from oracle import dgsct_oracle as O
from tests.golden_util import rel_l2

pytestmark = pytest.mark.gpu_experimental

SHAPES = {"vis_s0_small": (192, 2304, 96, 4096), "vis_s2": (768, 144, 384, 256), "aud_s3": (768, 64, 1536, 36)}


@pytest.mark.parametrize("flag", ["DGSCT_EARLY_RELEASE", "DGSCT_FUSE_BWD"])
@pytest.mark.parametrize("shape", list(SHAPES))
def test_flag_reproduces_default_path(flag, shape):
    C, N, Co, No = SHAPES[shape]
    cfg_o = O.make_config("ave", C, N, Co, No)
    cfg = dataclasses.asdict(cfg_o)
    p = {k: v.float().cuda().contiguous() for k, v in O.random_params(cfg_o, seed=41).items()}
    x, other = O.random_inputs(cfg_o, 4, seed=42)
    x, other = x.bfloat16().cuda(), other.bfloat16().cuda()
    g = torch.Generator().manual_seed(43)
    d_out = torch.randn(4, N, C, generator=g).bfloat16().cuda()
    d_smap = torch.randn(4, N, generator=g).bfloat16().cuda()

    def run():
        pp = {k: v.clone() for k, v in p.items()}
        out, smap, _ = adapter_forward_raw(cfg, pp, x, other, training=True)
        dx, do, gr = adapter_backward_raw(cfg, pp, x, other, d_out, d_smap, None, training=True)
        torch.cuda.synchronize()
        return out.float().cpu(), smap.float().cpu(), dx.float().cpu(), do.float().cpu(), {k: v.cpu() for k, v in gr.items()}

    ref = run()
    os.environ[flag] = "1"
    try:
        got = run()
    finally:
        os.environ.pop(flag, None)
    from dgsct_b200 import _lib as L
    assert L.lib().dgsct_tc_error_flag(1) == 0
    for a, b, name in zip(got[:4], ref[:4], ("out", "smap", "dx", "dother")):
        assert rel_l2(a, b) < 2e-2, (name, rel_l2(a, b))
    for k, v in ref[4].items():
        if float(v.abs().max()) > 1e-6 and k not in ("ln_before.bias",):
            assert rel_l2(got[4][k], v) < 5e-2, (k, rel_l2(got[4][k], v))


# ---- DGSCT_FC_FIRST changes the association of the resampler, so it is checked against the ORACLE (same bars as the
# default path: tests/test_gpu_parity.py, tests/test_gpu_backward.py), on the audio-direction shapes where it engages.
@pytest.mark.parametrize("shape", ["aud_s0", "aud_s1", "aud_s2", "aud_s3"])
def test_fc_first_forward_parity_bf16(shape, monkeypatch):
    from tests import test_gpu_parity as P
    monkeypatch.setenv("DGSCT_FC_FIRST", "1")
    P.test_real_shapes_bf16(shape)
    if shape in ("aud_s1", "aud_s3"):
        P.test_bf16_tensor_core_taps(shape)


@pytest.mark.parametrize("shape", ["aud_s1", "aud_s3"])
def test_fc_first_gradients_bf16(shape, monkeypatch):
    from tests import test_gpu_backward as B
    monkeypatch.setenv("DGSCT_FC_FIRST", "1")
    B.test_gradients_at_real_shapes(shape, torch.bfloat16)


def test_fc_first_train_mode_and_saved_activations(monkeypatch):
    from tests import test_gpu_parity as P
    from tests import test_gpu_backward as B
    monkeypatch.setenv("DGSCT_FC_FIRST", "1")
    P.test_train_mode_batch_stats("aud_s2", torch.bfloat16)
    B.test_saved_activation_backward_equals_recompute(torch.bfloat16)     # C=384 < Co=768: engages
