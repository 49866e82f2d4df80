"""What sits directly behind the adapters in the reference task networks, on the engine:

* :func:`pool_tokens` -- ``torch.bmm(spatial_att_maps, f)`` of DG-SCT/AVE/nets/net_trans.py:922,924: the LAST adapters'
  spatial softmax maps pool the token streams (this is what makes ``VisualAdapter``'s second return value load-bearing);
* :class:`TemporalGates` -- the sigmoid gates that close ``TemporalAttention.forward`` (net_trans.py:243-249; same idiom
  in the AVS-S4 copy, PVT_AVSModel.py:546-580).  Same sub-module names as the reference (``audio_gated.0``,
  ``video_gated.0``), so a reference checkpoint's ``temporal_attn.audio_gated.*`` entries load unchanged.

Both run in libdgsct.so (forward and backward); CUDA tensors only, no PyTorch fallback.  The LSTM / transformer part of
``TemporalAttention`` stays PyTorch (out of scope).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib as L


def _dt(t: torch.Tensor) -> int:
    if not t.is_cuda:
        raise RuntimeError("dgsct_b200 runs on CUDA tensors only (sm_100a); there is no CPU fallback")
    if t.dtype == torch.bfloat16:
        return L.DT_BF16
    if t.dtype == torch.float32:
        return L.DT_F32
    raise TypeError(f"float32 or bfloat16 expected, got {t.dtype}")


def _p(t):
    return C.c_void_p(t.data_ptr() if t is not None else 0)


class _PoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, smap, f):
        dt = _dt(f)
        if smap.dtype != f.dtype:
            raise TypeError("spatial_att_maps and f must have the same dtype")
        BT, N, Cc = f.shape
        m2 = smap.reshape(BT, N).contiguous()
        f2 = f.contiguous()
        out = torch.empty(BT, Cc, dtype=f.dtype, device=f.device)
        st = C.c_void_p(torch.cuda.current_stream(f.device).cuda_stream)
        with torch.cuda.device(f.device):
            L.check(L.lib().dgsct_pool_tokens_fwd(_p(m2), _p(f2), _p(out), BT, N, Cc, dt, st), "pool_tokens_fwd")
        ctx.save_for_backward(m2, f2)
        ctx.map_shape = smap.shape
        return out.view(BT, 1, Cc)

    @staticmethod
    def backward(ctx, d_out):
        m2, f2 = ctx.saved_tensors
        BT, N, Cc = f2.shape
        d_out = d_out.reshape(BT, Cc).to(f2.dtype).contiguous()
        need_m, need_f = ctx.needs_input_grad
        d_map = torch.empty_like(m2) if need_m else None
        d_f = torch.empty_like(f2) if need_f else None
        if need_m or need_f:
            st = C.c_void_p(torch.cuda.current_stream(f2.device).cuda_stream)
            with torch.cuda.device(f2.device):
                L.check(L.lib().dgsct_pool_tokens_bwd(_p(m2), _p(f2), _p(d_out), _p(d_map), _p(d_f), BT, N, Cc, _dt(f2), st),
                        "pool_tokens_bwd")
        return (d_map.view(ctx.map_shape) if need_m else None), d_f


def pool_tokens(spatial_att_maps: torch.Tensor, f: torch.Tensor) -> torch.Tensor:
    """Drop-in for ``torch.bmm(spatial_att_maps, f)`` at net_trans.py:922,924: ``spatial_att_maps`` (BT,1,N) is the second
    return value of the last adapter call, ``f`` (BT,N,C) the token stream; returns (BT,1,C)."""
    if f.dim() != 3 or spatial_att_maps.numel() != f.shape[0] * f.shape[1]:
        raise ValueError(f"expected spatial_att_maps (BT,1,N) and f (BT,N,C); got {tuple(spatial_att_maps.shape)} and {tuple(f.shape)}")
    return _PoolFn.apply(spatial_att_maps, f)


class _GateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a_kv, v_kv, vq, aq, wa, ba, wv, bv, gamma):
        dt = _dt(a_kv)
        shape = vq.shape
        D = shape[-1]
        rows = vq.numel() // D
        ts = [t.contiguous() for t in (a_kv, v_kv, vq, aq)]
        if any(t.dtype != a_kv.dtype or t.numel() != rows * D for t in ts):
            raise ValueError("a_kv, v_kv, vq and aq must share shape (T,B,D) and dtype")
        ws = [t.detach().reshape(-1).float().contiguous() for t in (wa, ba, wv, bv)]
        vq_out, aq_out = torch.empty_like(ts[2]), torch.empty_like(ts[3])
        av = torch.empty(*shape[:-1], 1, dtype=vq.dtype, device=vq.device)
        gates = torch.empty(rows, 2, dtype=torch.float32, device=vq.device)
        st = C.c_void_p(torch.cuda.current_stream(vq.device).cuda_stream)
        with torch.cuda.device(vq.device):
            L.check(L.lib().dgsct_temporal_gates_fwd(*[_p(t) for t in ts], *[_p(w) for w in ws], C.c_float(gamma), _p(vq_out), _p(aq_out),
                                                     _p(av), _p(gates), rows, D, dt, st), "temporal_gates_fwd")
        ctx.save_for_backward(*ts, ws[0], ws[2], gates)
        ctx.gamma, ctx.wshapes = gamma, (wa.shape, ba.shape, wv.shape, bv.shape)
        return vq_out, aq_out, av

    @staticmethod
    def backward(ctx, d_vq_out, d_aq_out, d_av):
        a_kv, v_kv, vq, aq, wa, wv, gates = ctx.saved_tensors
        D = vq.shape[-1]
        rows = vq.numel() // D
        dt = _dt(vq)
        d_vq_out = (d_vq_out if d_vq_out is not None else torch.zeros_like(vq)).to(vq.dtype).contiguous()
        d_aq_out = (d_aq_out if d_aq_out is not None else torch.zeros_like(aq)).to(vq.dtype).contiguous()
        d_av = d_av.to(vq.dtype).contiguous() if d_av is not None else None
        outs = [torch.empty_like(t) for t in (a_kv, v_kv, vq, aq)]
        d_wa, d_wv = torch.zeros(D, dtype=torch.float32, device=vq.device), torch.zeros(D, dtype=torch.float32, device=vq.device)
        d_ba, d_bv = torch.zeros(1, dtype=torch.float32, device=vq.device), torch.zeros(1, dtype=torch.float32, device=vq.device)
        st = C.c_void_p(torch.cuda.current_stream(vq.device).cuda_stream)
        with torch.cuda.device(vq.device):
            L.check(L.lib().dgsct_temporal_gates_bwd(_p(a_kv), _p(v_kv), _p(vq), _p(aq), _p(wa), _p(wv), C.c_float(ctx.gamma), _p(gates),
                                                     _p(d_vq_out), _p(d_aq_out), _p(d_av), *[_p(o) for o in outs], _p(d_wa), _p(d_ba),
                                                     _p(d_wv), _p(d_bv), rows, D, dt, st), "temporal_gates_bwd")
        s = ctx.wshapes
        return (*outs, d_wa.view(s[0]), d_ba.view(s[1]), d_wv.view(s[2]), d_bv.view(s[3]), None)


class TemporalGates(nn.Module):
    """The gate tail of the reference ``TemporalAttention`` (net_trans.py:199-213, 243-249).  Replace, in its ``forward``::

        audio_gate = self.audio_gated(audio_key_value_feature) ... return video_query_output, audio_query_output, audio_visual_gate

    by ``return self.gates(audio_key_value_feature, video_key_value_feature, video_query_output, audio_query_output)``
    (or bind an existing module's two heads with :meth:`from_reference`)."""

    def __init__(self, d_model: int = 256, gamma: float = 0.1):
        super().__init__()
        self.audio_gated = nn.Sequential(nn.Linear(d_model, 1), nn.Sigmoid())
        self.video_gated = nn.Sequential(nn.Linear(d_model, 1), nn.Sigmoid())
        self.gamma = float(gamma)

    @classmethod
    def from_reference(cls, temporal_attention: nn.Module) -> "TemporalGates":
        """Share (not copy) the gate parameters of a reference ``TemporalAttention`` instance."""
        g = cls.__new__(cls)
        nn.Module.__init__(g)
        g.audio_gated, g.video_gated = temporal_attention.audio_gated, temporal_attention.video_gated
        g.gamma = float(temporal_attention.gamma)
        return g

    def forward(self, audio_key_value_feature, video_key_value_feature, video_query_output, audio_query_output):
        la, lv = self.audio_gated[0], self.video_gated[0]
        return _GateFn.apply(audio_key_value_feature, video_key_value_feature, video_query_output, audio_query_output,
                             la.weight, la.bias, lv.weight, lv.bias, self.gamma)
