"""Parity at BASELINE.json's full size (B=8 clips x T=10 frames = 80 frames per call) through size-independent
properties: every adapter op is per-frame in eval mode, so

* the 80-frame call must reproduce 10-frame calls on slices of the same inputs (forward outputs, dx, dother);
* parameter gradients are additive over frames: the 80-frame gradient is the sum of the per-clip gradients
  (a checksum of checksums);
* spatial_att_maps rows sum to 1;
* one clip of the 80-frame output is checked against the CPU oracle under the north-star's bf16 bar.

The two paths run the same kernels; they differ only in tiling (frames flattened into the M extent, frames accumulated
per tile in the weight-gradient products) and in the order of the fp32 atomics."""
import pytest
import torch

from dgsct_b200.adapter import adapter_backward_raw, adapter_forward_raw
from oracle import dgsct_oracle as O
from tests.golden_util import rel_l2
from tests.test_gpu_parity import SHAPES, cfg_dict, to_dev

pytestmark = pytest.mark.gpu
BT, CLIP = 80, 10


def _inputs(shape, seed):
    C, N, Co, No = SHAPES[shape]
    cfg = O.make_config("ave", C, N, Co, No)
    params = O.random_params(cfg, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(BT, N, C, generator=g).bfloat16()
    other = torch.randn(BT, No, Co, generator=g).bfloat16()
    return cfg, params, x, other


@pytest.mark.parametrize("shape", ["vis_s0", "aud_s0", "vis_s2", "aud_s3"])
def test_forward_at_full_batch(shape):
    cfg, params, x, other = _inputs(shape, 51)
    p = to_dev(params)
    xd, od = x.cuda(), other.cuda()
    out, smap, _ = adapter_forward_raw(cfg_dict(cfg), p, xd, od, training=False)
    torch.cuda.synchronize()
    from dgsct_b200 import _lib as L
    assert L.lib().dgsct_tc_error_flag(1) == 0
    assert float((smap.float().sum(-1) - 1).abs().max()) < 2e-2
    for lo in (0, BT - CLIP):
        o2, s2, _ = adapter_forward_raw(cfg_dict(cfg), p, xd[lo:lo + CLIP].contiguous(), od[lo:lo + CLIP].contiguous(), training=False)
        assert rel_l2(out[lo:lo + CLIP].float().cpu(), o2.float().cpu()) < 2e-3, (shape, lo)
        assert rel_l2(smap[lo:lo + CLIP].float().cpu(), s2.float().cpu()) < 2e-3, (shape, lo)
    lo = BT - CLIP
    ro, rs, _, _ = O.adapter_forward(cfg, params, x[lo:].float(), other[lo:].float(), training=False)
    assert rel_l2(out[lo:].float().cpu(), ro) < 1e-2
    assert rel_l2(smap[lo:].float().cpu(), rs.reshape(CLIP, -1)) < 1e-2


@pytest.mark.parametrize("shape", ["vis_s1", "aud_s2"])
def test_backward_at_full_batch_is_the_sum_of_per_clip_backwards(shape):
    cfg, params, x, other = _inputs(shape, 61)
    C, N, Co, No = SHAPES[shape]
    p = to_dev(params)
    xd, od = x.cuda(), other.cuda()
    g = torch.Generator().manual_seed(62)
    d_out = (torch.randn(BT, N, C, generator=g) * 1e-1).bfloat16().cuda()
    d_smap = torch.randn(BT, N, generator=g).bfloat16().cuda()
    dx, do, grads = adapter_backward_raw(cfg_dict(cfg), p, xd, od, d_out, d_smap, None, training=False)
    torch.cuda.synchronize()
    acc = {k: torch.zeros_like(v) for k, v in grads.items()}
    for lo in range(0, BT, CLIP):
        sl = slice(lo, lo + CLIP)
        dx2, do2, g2 = adapter_backward_raw(cfg_dict(cfg), p, xd[sl].contiguous(), od[sl].contiguous(), d_out[sl].contiguous(),
                                            d_smap[sl].contiguous(), None, training=False)
        if lo in (0, BT - CLIP):
            assert rel_l2(dx[sl].float().cpu(), dx2.float().cpu()) < 5e-3, (shape, lo)
            assert rel_l2(do[sl].float().cpu(), do2.float().cpu()) < 5e-3, (shape, lo)
        for k, v in g2.items():
            acc[k] += v
    torch.cuda.synchronize()
    bad = {}
    for k, v in grads.items():
        if float(acc[k].abs().max()) > 1e-6:
            e = rel_l2(v.cpu(), acc[k].cpu())
            if e > 2e-2:
                bad[k] = f"{e:.2e}"
    assert not bad, bad
