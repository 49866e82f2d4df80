"""Host-side mirror of the reference's ``VisualAdapter`` (DG-SCT/AVE/nets/net_trans.py:433-674 and the
per-task copies) on top of the C ABI in ``include/dgsct.h``.

* same constructor signature, parameter/buffer names and shapes (checkpoints load unchanged,
  main_trans.py:298,306), same ``forward(x, vis_token)`` contract: ``x`` is ``(BT, C, N, 1)``,
  ``vis_token`` is ``(BT, C', N', 1)``; returns ``(output (BT,C,N,1), spatial_att_maps (BT,1,N))`` and, for
  the few-shot / pretrain / zero-shot variants, ``temporal_att_maps (B,T,1,1)`` as a third value.
* all arithmetic runs in libdgsct.so on the current CUDA stream; PyTorch only owns memory.  There is no
  CPU path and no PyTorch fallback: CPU tensors or a missing library raise.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib as L

# variant -> (alpha, beta, gamma, resampler, ln_before allowed, gate_before_ln, temporal returned, T)
VARIANT_DEFAULTS = {
    "ave":      dict(alpha=0.3, beta=0.05, gamma=0.0, resampler="conv", honour_ln_before=True, gate_before_ln=False, temporal=False, T=10),
    "avvp":     dict(alpha=0.3, beta=0.05, gamma=0.0, resampler="conv", honour_ln_before=True, gate_before_ln=False, temporal=False, T=10),
    "avqa":     dict(alpha=0.3, beta=0.05, gamma=0.0, resampler="conv", honour_ln_before=True, gate_before_ln=False, temporal=False, T=10),
    "s4":       dict(alpha=0.3, beta=0.05, gamma=0.0, resampler="bicubic", honour_ln_before=False, gate_before_ln=True, temporal=False, T=5),
    "ms3":      dict(alpha=0.2, beta=0.1, gamma=0.0, resampler="conv", honour_ln_before=False, gate_before_ln=True, temporal=False, T=5),
    "pretrain": dict(alpha=0.3, beta=0.01, gamma=0.05, resampler="conv", honour_ln_before=True, gate_before_ln=False, temporal=True, T=10),
    "zeroshot": dict(alpha=0.3, beta=0.01, gamma=0.05, resampler="conv", honour_ln_before=True, gate_before_ln=False, temporal=True, T=10),
    "fewshot":  dict(alpha=None, beta=None, gamma=None, resampler="conv", honour_ln_before=True, gate_before_ln=False, temporal=True, T=10),
}

# state_dict name -> field of dgsct_params
_PARAM_MAP = {
    "conv_adapter.weight": "resample_w", "conv_adapter.bias": "resample_b", "fc.weight": "fc_w", "fc.bias": "fc_b",
    "fc_affine_audio_1.weight": "a1_w", "fc_affine_audio_1.bias": "a1_b",
    "fc_affine_video_1.weight": "v1_w", "fc_affine_video_1.bias": "v1_b",
    "fc_affine_bottleneck.weight": "bott_w", "fc_affine_bottleneck.bias": "bott_b",
    "fc_affine_video_2.weight": "v2_w", "fc_affine_video_2.bias": "v2_b",
    "fc_affine_audio_2.weight": "a2_w", "fc_affine_audio_2.bias": "a2_b",
    "fc_affine_v_s_att.weight": "s_w", "fc_affine_v_s_att.bias": "s_b",
    "fc_affine_v_c_att.weight": "c_w", "fc_affine_v_c_att.bias": "c_b",
    "my_tokens": "tokens", "gate_av": "gate_av", "gate": "gate",
    "down_sampler.weight": "down_w", "up_sampler.weight": "up_w",
    "bn1.weight": "bn1_w", "bn1.bias": "bn1_b", "bn2.weight": "bn2_w", "bn2.bias": "bn2_b",
    "bn1.running_mean": "bn1_rm", "bn1.running_var": "bn1_rv", "bn2.running_mean": "bn2_rm", "bn2.running_var": "bn2_rv",
    "ln_before.weight": "lnb_w", "ln_before.bias": "lnb_b", "ln_post.weight": "lnp_w", "ln_post.bias": "lnp_b",
    "temporal_gated.0.weight": "t_w", "temporal_gated.0.bias": "t_b",
}

_workspaces: Dict[tuple, torch.Tensor] = {}
_captured_keys: set = set()          # workspaces a CUDA graph was captured against
_retired: list = []                  # ... are never freed: graph replays keep writing into them


def _workspace(device: torch.device, stream_ptr, nbytes: int) -> torch.Tensor:
    """Scratch cached per (device, stream).  A workspace that was handed out while the stream was being captured into a
    CUDA graph is pinned for the life of the process: when a later, larger call regrows the cache entry the old tensor
    moves to ``_retired`` instead of being freed, so replays of the captured graph never write into recycled memory."""
    key = (device.index, stream_ptr)
    ws = _workspaces.get(key)
    capturing = torch.cuda.is_current_stream_capturing()
    if ws is None or ws.numel() < nbytes:
        if ws is not None and key in _captured_keys:
            _retired.append(ws)
            _captured_keys.discard(key)
        ws = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    if capturing:
        _captured_keys.add(key)
    return ws


_bicubic_cache: Dict[tuple, torch.Tensor] = {}


def bicubic_matrix(n_in_side: int, n_out_side: int, device) -> torch.Tensor:
    """(N, No) fp32 interpolation matrix of the AVS-S4 resampler (PVT_AVSModel.py:190-197), built by the
    engine's host helper and cached per device."""
    key = (n_in_side, n_out_side, str(device))
    w = _bicubic_cache.get(key)
    if w is None:
        host = torch.empty(n_out_side * n_out_side, n_in_side * n_in_side, dtype=torch.float32)
        L.check(L.lib().dgsct_bicubic_matrix(n_in_side, n_out_side, C.c_void_p(host.data_ptr())), "bicubic_matrix")
        w = host.to(device)
        _bicubic_cache[key] = w
    return w


_bicubic_csr_cache: Dict[tuple, tuple] = {}


def bicubic_csr(n_in_side: int, n_out_side: int, device):
    """CSR of the AVS-S4 bicubic resampler (PVT_AVSModel.py:190-197) and of its transpose, built by the engine's host
    helper and cached per device: ((rowptr, col, val), (rowptr_T, col_T, val_T)) -- int32 / fp32 device tensors."""
    key = (n_in_side, n_out_side, str(device))
    hit = _bicubic_csr_cache.get(key)
    if hit is None:
        lib = L.lib()
        cap = lib.dgsct_bicubic_csr_capacity(n_in_side, n_out_side)
        out = []
        for transpose, rows in ((0, n_out_side * n_out_side), (1, n_in_side * n_in_side)):
            rp = torch.empty(rows + 1, dtype=torch.int32)
            ci = torch.empty(cap, dtype=torch.int32)
            va = torch.empty(cap, dtype=torch.float32)
            nnz = lib.dgsct_bicubic_csr(n_in_side, n_out_side, transpose, C.c_void_p(rp.data_ptr()), C.c_void_p(ci.data_ptr()),
                                        C.c_void_p(va.data_ptr()))
            if nnz < 0:
                L.check(nnz, "bicubic_csr")
            out.append((rp.to(device), ci[:nnz].to(device), va[:nnz].to(device)))
        hit = tuple(out)
        _bicubic_csr_cache[key] = hit
    return hit


def _describe(cfg: dict, params: Dict[str, torch.Tensor], x: torch.Tensor, other: torch.Tensor, training: bool,
              force_simt: bool = False):
    """Validate the call and build the C-ABI descriptor + parameter table.  Returns (desc, params_struct, keep, x, other)."""
    if not (x.is_cuda and other.is_cuda):
        raise RuntimeError("dgsct_b200 runs on CUDA tensors only (sm_100a); there is no CPU fallback")
    if x.dtype != other.dtype or x.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError(f"x and vis_token must both be float32 or bfloat16 (got {x.dtype}, {other.dtype})")
    if x.dim() != 3 or other.dim() != 3 or x.shape[0] != other.shape[0]:
        raise ValueError(f"expected x (BT,N,C) and other (BT,No,Co); got {tuple(x.shape)} and {tuple(other.shape)}")
    if not x.is_contiguous():
        x = x.contiguous()
    if not other.is_contiguous():
        other = other.contiguous()
    BT, N, Cc = x.shape
    _, No, Co = other.shape
    dev = x.device
    d = L.Desc()
    d.BT, d.T, d.C, d.N, d.Co, d.No = BT, int(cfg.get("T", 10)), Cc, N, Co, No
    d.K, d.r, d.g = int(cfg["K"]), int(cfg["r"]), int(cfg["g"])
    d.io_dtype = L.DT_BF16 if x.dtype == torch.bfloat16 else L.DT_F32
    d.alpha, d.beta, d.gamma = float(cfg["alpha"]), float(cfg["beta"]), float(cfg.get("gamma", 0.0))
    d.eps_ln, d.eps_bn, d.bn_momentum = float(cfg.get("eps_ln", 1e-5)), float(cfg.get("eps_bn", 1e-5)), float(cfg.get("bn_momentum", 0.1))
    flags = 0
    flags |= L.F_TRAIN if training else 0
    flags |= L.F_USE_BN if cfg.get("use_bn", True) else 0
    flags |= L.F_USE_GATE if cfg.get("use_gate", True) else 0
    flags |= L.F_LN_BEFORE if cfg.get("ln_before", True) else 0
    flags |= L.F_LN_POST if cfg.get("ln_post", True) else 0
    flags |= L.F_GATE_BEFORE_LN if cfg.get("gate_before_ln", False) else 0
    flags |= L.F_TEMPORAL if cfg.get("temporal", False) else 0
    flags |= L.F_FORCE_SIMT if force_simt else 0
    keep = []   # keep temporaries alive until the launch calls return
    P = L.Params()
    bicubic = cfg.get("resampler", "conv") == "bicubic"
    for name, field in _PARAM_MAP.items():
        t = params.get(name)
        if bicubic and field in ("resample_w", "resample_b"):
            continue
        if t is None:
            continue
        if t.dtype != torch.float32 or not t.is_cuda or t.device != dev:
            raise TypeError(f"parameter {name} must be a float32 CUDA tensor on {dev} (got {t.dtype} on {t.device})")
        if not t.is_contiguous():
            t = t.contiguous()
            keep.append(t)
        setattr(P, field, t.data_ptr())
    if bicubic:
        s_in, s_out = int(math.isqrt(No)), int(math.isqrt(N))
        if s_in * s_in != No or s_out * s_out != N:
            raise ValueError("the bicubic resampler needs square token grids")
        flags |= L.F_NO_CONV_BIAS
        if cfg.get("dense_bicubic", False) or Co % 8 or Cc % 8:
            # dense (N, No) interpolation matrix through the GEMM path (bf16 coefficients on the tensor-core path)
            P.resample_w = bicubic_matrix(s_in, s_out, dev).data_ptr()
        else:
            # 16-tap stencil as CSR: one gather kernel, fp32 coefficients (forward) and its transpose (backward)
            (rp, ci, va), (rpt, cit, vat) = bicubic_csr(s_in, s_out, dev)
            P.rs_rowptr, P.rs_col, P.rs_val = rp.data_ptr(), ci.data_ptr(), va.data_ptr()
            P.rst_rowptr, P.rst_col, P.rst_val = rpt.data_ptr(), cit.data_ptr(), vat.data_ptr()
    else:
        w = params.get("conv_adapter.weight")
        if w is None or w.numel() != N * No:
            raise ValueError(f"conv_adapter.weight must have {N}x{No} elements (conv_dim_out x conv_dim_in)")
    if tuple(params["fc.weight"].shape) != (Cc, Co):
        raise ValueError(f"fc.weight is {tuple(params['fc.weight'].shape)}, inputs need ({Cc},{Co})")
    if params["my_tokens"].shape[0] != d.K:
        raise ValueError("my_tokens rows != K")
    d.flags = flags
    return d, P, keep, x, other


def adapter_forward_raw(cfg: dict, params: Dict[str, torch.Tensor], x: torch.Tensor, other: torch.Tensor,
                        training: bool = False, want_taps: bool = False, force_simt: bool = False,
                        packed: Optional[torch.Tensor] = None, pack_only: bool = False,
                        keep: Optional[list] = None):
    """Token-major entry: x (BT,N,C), other (BT,No,Co) contiguous CUDA tensors (fp32 or bf16); ``params`` maps
    reference state_dict names to fp32 CUDA tensors.  ``cfg`` keys: K, r, g, T, alpha, beta, gamma, resampler,
    ln_before, ln_post, gate_before_ln, use_bn, use_gate, temporal, eps_ln, eps_bn, bn_momentum.
    Returns (out (BT,N,C), smap (BT,N), tmap (BT,) fp32 | None[, taps dict]).
    ``keep``: pass an empty list to run the forward with DGSCT_F_KEEP into a workspace of its own (every intermediate
    the backward needs survives the call); the list receives that workspace tensor, to be handed to
    ``adapter_backward_raw(saved=...)``."""
    d, P, _alive, x, other = _describe(cfg, params, x, other, training, force_simt)
    if keep is not None:
        d.flags |= L.F_KEEP
    BT, N, Cc = x.shape
    dev = x.device
    lib = L.lib()
    nbytes = lib.dgsct_workspace_bytes(C.byref(d))
    if nbytes == 0:
        raise ValueError(f"dgsct: invalid adapter description C={Cc} N={N} Co={d.Co} No={d.No} K={d.K} r={d.r} g={d.g} T={d.T} BT={BT}")
    stream = torch.cuda.current_stream(dev).cuda_stream
    if pack_only:
        pb = lib.dgsct_packed_bytes(C.byref(d))
        buf = torch.empty(pb + 256, dtype=torch.uint8, device=dev)
        off = (-buf.data_ptr()) % 256
        with torch.cuda.device(dev):
            L.check(lib.dgsct_pack_params(C.byref(d), C.byref(P), C.c_void_p(buf.data_ptr() + off), C.c_void_p(stream)), "pack_params")
        return buf[off:off + pb]
    if keep is not None:
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)      # owned by the caller (autograd ctx)
        keep.append(ws)
    else:
        ws = _workspace(dev, stream, nbytes + 256)
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    out = torch.empty_like(x)
    smap = torch.empty(BT, N, dtype=x.dtype, device=dev)
    tmap = torch.empty(BT, dtype=torch.float32, device=dev) if cfg.get("temporal", False) else None
    taps = None
    taps_t = None
    if want_taps:
        Cd = Cc // d.r
        taps_t = dict(V=torch.empty(BT, N, Cc, device=dev), Tp=torch.empty(BT, d.K, Cc, device=dev),
                      X1=torch.empty(BT, N, Cc, device=dev), obar=torch.empty(BT, Cc, device=dev),
                      m=torch.empty(BT, Cc, device=dev), s=torch.empty(BT, N, device=dev),
                      Y=torch.empty(BT, N, Cc, device=dev), Z=torch.empty(BT, N, Cd, device=dev))
        taps = L.Taps()
        for k, v in taps_t.items():
            setattr(taps, k, v.data_ptr())
    with torch.cuda.device(dev):
        rc = lib.dgsct_adapter_fwd(C.byref(d), C.byref(P), C.c_void_p(packed.data_ptr() if packed is not None else 0),
                                   C.c_void_p(x.data_ptr()), C.c_void_p(other.data_ptr()),
                                   C.c_void_p(out.data_ptr()), C.c_void_p(smap.data_ptr()),
                                   C.c_void_p(tmap.data_ptr() if tmap is not None else 0),
                                   C.byref(taps) if taps is not None else None,
                                   C.c_void_p(ws_ptr), C.c_size_t(nbytes), C.c_void_p(stream))
    L.check(rc, "adapter_fwd")
    if want_taps:
        taps_t["obar"] = taps_t["obar"] / N
        return out, smap, tmap, taps_t
    return out, smap, tmap


# reverse of _PARAM_MAP for the gradient table (running stats have no gradient)
_GRAD_NAMES = {v: k for k, v in _PARAM_MAP.items() if "running" not in k}


def adapter_backward_raw(cfg: dict, params: Dict[str, torch.Tensor], x: torch.Tensor, other: torch.Tensor,
                         d_out: torch.Tensor, d_smap: Optional[torch.Tensor] = None, d_tmap: Optional[torch.Tensor] = None,
                         training: bool = True, packed: Optional[torch.Tensor] = None,
                         saved: Optional[torch.Tensor] = None, grad_targets: Optional[Dict[str, torch.Tensor]] = None):
    """Gradients of the adapter w.r.t. x, other and every parameter.  With ``saved`` (the workspace kept by a
    ``keep=[]`` forward) the call is the backward proper (dgsct_adapter_bwd_saved); without it the library recomputes
    the forward inside the call (dgsct_adapter_bwd).  The C ABI ADDS parameter gradients into the buffers it is given:
    names present in ``grad_targets`` (contiguous fp32 tensors, e.g. ``param.grad`` views of a gradient bucket) are
    accumulated in place; every other gradient lands in one freshly zeroed flat buffer.
    Returns (dx, dother, {state_dict name: fp32 grad} for the names NOT in grad_targets)."""
    d, P, _alive, x, other = _describe(cfg, params, x, other, training)
    dev = x.device
    lib = L.lib()
    d_out = d_out.to(x.dtype).contiguous()
    if d_smap is not None:
        d_smap = d_smap.to(x.dtype).contiguous()
    if d_tmap is not None:
        d_tmap = d_tmap.to(torch.float32).contiguous()
    G = L.Grads()
    grads: Dict[str, torch.Tensor] = {}
    bicubic = cfg.get("resampler", "conv") == "bicubic"
    fresh = []
    for field, name in _GRAD_NAMES.items():
        t = params.get(name)
        if t is None or (bicubic and field in ("resample_w", "resample_b")):
            continue
        if field in ("t_w", "t_b") and not cfg.get("temporal", False):
            continue
        tgt = grad_targets.get(name) if grad_targets else None
        if tgt is not None:
            if tgt.dtype != torch.float32 or not tgt.is_contiguous() or tgt.numel() != t.numel() or tgt.device != dev:
                raise TypeError(f"gradient target for {name} must be a contiguous float32 tensor of {t.numel()} elements on {dev}")
            setattr(G, field, tgt.data_ptr())
        else:
            fresh.append((field, name, t))
    if fresh:
        offs, total = [], 0
        for _, _, t in fresh:
            offs.append(total)
            total += (t.numel() + 3) // 4 * 4               # keep every view 16-byte aligned
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        for (field, name, t), o in zip(fresh, offs):
            g = flat[o:o + t.numel()].view(t.shape)
            grads[name] = g
            setattr(G, field, g.data_ptr())
    stream = torch.cuda.current_stream(dev).cuda_stream
    if saved is not None:
        nbytes = lib.dgsct_bwd_saved_workspace_bytes(C.byref(d))
    else:
        nbytes = lib.dgsct_bwd_workspace_bytes(C.byref(d))
    if nbytes == 0:
        raise ValueError("dgsct: invalid adapter description for backward")
    ws = _workspace(dev, ("bwd", stream), nbytes + 256)
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    dx = torch.empty_like(x)
    dother = torch.empty_like(other)
    pk = C.c_void_p(packed.data_ptr() if packed is not None else 0)
    tail = (C.c_void_p(x.data_ptr()), C.c_void_p(other.data_ptr()), C.c_void_p(d_out.data_ptr()),
            C.c_void_p(d_smap.data_ptr() if d_smap is not None else 0),
            C.c_void_p(d_tmap.data_ptr() if d_tmap is not None else 0),
            C.c_void_p(dx.data_ptr()), C.c_void_p(dother.data_ptr()), C.byref(G),
            C.c_void_p(ws_ptr), C.c_size_t(nbytes), C.c_void_p(stream))
    with torch.cuda.device(dev):
        if saved is not None:
            saved_ptr = (saved.data_ptr() + 255) // 256 * 256
            rc = lib.dgsct_adapter_bwd_saved(C.byref(d), C.byref(P), pk, C.c_void_p(saved_ptr), *tail)
        else:
            rc = lib.dgsct_adapter_bwd(C.byref(d), C.byref(P), pk, *tail)
    L.check(rc, "adapter_bwd")
    return dx, dother, grads


class _AdapterFn(torch.autograd.Function):
    """autograd bridge: forward and backward both run in libdgsct.so (no PyTorch autograd fallback).
    ``module.save_activations`` (default True): the training forward keeps its intermediates in a workspace owned by
    the autograd context and the backward runs from them; False: only the inputs are saved and the backward call
    recomputes the forward (activation recomputation, for memory-constrained runs).
    ``module.fused_grad_accumulation`` (default False): parameter gradients are accumulated by the kernels straight
    into existing contiguous fp32 ``param.grad`` tensors (e.g. views of a GradBucketer bucket) and autograd receives
    None for them -- no temporary gradient tensors, no AccumulateGrad add."""

    @staticmethod
    def forward(ctx, module, names, xt, ot, *flat_params):
        cfg, params = module._cfg(), module._param_dict()
        packed = module._packed_weights(cfg, params, xt, ot)
        keep = [] if getattr(module, "save_activations", True) else None
        out, smap, tmap = adapter_forward_raw(cfg, params, xt, ot, training=module.training, packed=packed, keep=keep)
        ctx.module, ctx.names, ctx.training, ctx.packed = module, names, module.training, packed
        ctx.saved_ws = keep[0] if keep else None
        ctx.save_for_backward(xt, ot)
        if tmap is not None:
            return out, smap, tmap
        return out, smap

    @staticmethod
    def backward(ctx, *gouts):
        xt, ot = ctx.saved_tensors
        module = ctx.module
        cfg, params = module._cfg(), module._param_dict()
        d_out = gouts[0]
        if d_out is None:
            d_out = torch.zeros_like(xt)
        d_smap = gouts[1] if len(gouts) > 1 else None
        d_tmap = gouts[2] if len(gouts) > 2 else None
        targets = None
        if getattr(module, "fused_grad_accumulation", False):
            targets = {}
            for n, p in module._named_tensors():
                g = p.grad if p.is_leaf else None
                if n in ctx.names and g is not None and g.dtype == torch.float32 and g.is_contiguous() and g.device == xt.device:
                    targets[n] = g
        dx, dother, grads = adapter_backward_raw(cfg, params, xt, ot, d_out, d_smap, d_tmap, training=ctx.training,
                                                 packed=ctx.packed, saved=ctx.saved_ws, grad_targets=targets)
        ctx.saved_ws = None
        flat = []
        for n in ctx.names:
            g = grads.get(n)
            p = params.get(n)
            flat.append(g.view_as(p) if (g is not None and p is not None) else None)
        # AdapterDDP: every kernel that adds into this module's bucket views has now been enqueued on the current stream
        # (a module full-backward hook would fire BEFORE this function when no input requires grad)
        cb = getattr(module, "_ddp_done_cb", None)
        if cb is not None:
            cb()
        return (None, None, dx, dother, *flat)


class VisualAdapter(nn.Module):
    """Drop-in replacement of the reference class.  ``variant`` selects which task tree's semantics to follow
    (the reference keeps one near-identical copy of the class per task)."""

    def __init__(self, input_dim, output_dim, adapter_kind, dim_list=None, layer_idx=0, reduction_factor=16, opt=None,
                 use_bn=True, use_gate=True, num_tk=87, conv_dim_in=0, conv_dim_out=0, linear_in=0, linear_out=0,
                 variant="ave"):
        super().__init__()
        if adapter_kind != "bottleneck" or not getattr(opt, "is_multimodal", 1):
            # the "basic" and non-multimodal branches are dead in every task
            raise NotImplementedError("dgsct_b200 implements adapter_kind='bottleneck' with opt.is_multimodal=1")
        if variant not in VARIANT_DEFAULTS:
            raise ValueError(f"unknown variant {variant!r}")
        if input_dim != output_dim or linear_out != input_dim:
            raise ValueError("input_dim, output_dim and linear_out must be equal")
        self.variant = variant
        self.adapter_kind = adapter_kind
        self.use_bn = use_bn
        self.is_multimodal = opt.is_multimodal
        self.opt = opt
        # AVQA / AVS copies take K from opt.num_tokens and have no num_tk argument (net_avst.py:31,60)
        self.num_tk = int(getattr(opt, "num_tokens", num_tk)) if variant in ("avqa", "s4", "ms3") else int(num_tk)
        self.reduction_factor = int(reduction_factor)
        self.conv_dim_in, self.conv_dim_out = conv_dim_in, conv_dim_out
        self.conv_adapter = nn.Conv2d(conv_dim_in, conv_dim_out, kernel_size=1)
        self.fc = nn.Linear(linear_in, linear_out)
        d_model = linear_out // 2
        self.fc_affine_audio_1 = nn.Linear(linear_out, linear_out)
        self.fc_affine_video_1 = nn.Linear(linear_out, linear_out)
        self.fc_affine_bottleneck = nn.Linear(linear_out, d_model)
        self.fc_affine_video_2 = nn.Linear(linear_out, d_model)
        self.fc_affine_audio_2 = nn.Linear(linear_out, d_model)
        self.fc_affine_v_s_att = nn.Linear(d_model, 1)
        self.fc_affine_v_c_att = nn.Linear(d_model, linear_out)
        if variant in ("avvp", "pretrain", "zeroshot", "fewshot"):
            self.fc_caption = nn.Linear(512, 192)      # dead in the reference (caption is never passed)
        if variant in ("avvp", "s4", "ms3", "pretrain", "zeroshot", "fewshot"):
            self.temporal_gated = nn.Sequential(nn.Linear(linear_out, 1), nn.Sigmoid())
        self.gate = nn.Parameter(torch.zeros(1)) if use_gate else None
        self.down_sample_size = input_dim // reduction_factor
        if variant in ("avqa", "s4", "ms3"):
            self.my_tokens = nn.Parameter(torch.zeros((self.num_tk, input_dim)))
        else:
            self.my_tokens = nn.Parameter(torch.rand((self.num_tk, input_dim)))
            self.gate_tk = nn.Parameter(torch.ones(1))   # declared, never used (net_trans.py:487)
        self.gate_av = nn.Parameter(torch.zeros(1))
        g = int(opt.num_conv_group)
        self.down_sampler = nn.Conv2d(input_dim, self.down_sample_size, 1, groups=g, bias=False)
        self.up_sampler = nn.Conv2d(self.down_sample_size, output_dim, 1, groups=g, bias=False)
        if use_bn:
            self.bn1 = nn.BatchNorm2d(self.down_sample_size)
            self.bn2 = nn.BatchNorm2d(output_dim)
        if getattr(opt, "is_before_layernorm", 0):
            self.ln_before = nn.LayerNorm(output_dim)
        if getattr(opt, "is_post_layernorm", 0):
            self.ln_post = nn.LayerNorm(output_dim)

    # -- engine description ---------------------------------------------------------------------------
    def _cfg(self) -> dict:
        v = VARIANT_DEFAULTS[self.variant]
        alpha, beta, gamma = v["alpha"], v["beta"], v["gamma"]
        if alpha is None:   # few-shot reads them from the CLI options (few-shot/nets/net_trans.py:537,541)
            alpha, beta, gamma = float(self.opt.alpha), float(self.opt.beta), float(self.opt.gamma)
        return dict(K=self.num_tk, r=self.reduction_factor, g=int(self.opt.num_conv_group), T=v["T"],
                    alpha=alpha, beta=beta, gamma=gamma, resampler=v["resampler"],
                    ln_before=bool(getattr(self.opt, "is_before_layernorm", 0)) and v["honour_ln_before"],
                    ln_post=bool(getattr(self.opt, "is_post_layernorm", 0)),
                    gate_before_ln=v["gate_before_ln"], use_bn=bool(self.use_bn), use_gate=self.gate is not None,
                    temporal=v["temporal"], eps_ln=1e-5, eps_bn=1e-5, bn_momentum=0.1)

    def _named_tensors(self):
        """(state_dict name, tensor) for every engine parameter / buffer, resolved by ATTRIBUTE walk rather than through
        ``named_parameters()``: ``nn.DataParallel`` replicas (S4 / MS3 / AVQA scripts, avs_s4/train.py:139) keep their
        broadcast copies as plain attributes and have empty ``_parameters`` dicts."""
        out = []
        for name in _PARAM_MAP:
            obj = self
            for part in name.split("."):
                if isinstance(obj, nn.Sequential) and part.isdigit():
                    obj = obj[int(part)] if int(part) < len(obj) else None
                else:
                    obj = getattr(obj, part, None)
                if obj is None:
                    break
            if isinstance(obj, torch.Tensor):
                out.append((name, obj))
        return out

    def _param_dict(self) -> Dict[str, torch.Tensor]:
        return {k: v.detach() for k, v in self._named_tensors()}

    def invalidate_packed(self) -> None:
        """Drop the cached bf16 operand copies.  The cache is keyed on the parameters' ``_version``, which updates made
        through ``.data`` (manual ``p.data.copy_``, EMA, ``dist.broadcast(p.data)``) do not bump: call this after them."""
        self._packed_key = None
        self._packed_buf = None

    _PACKED_SOURCES = ("conv_adapter.weight", "fc.weight", "my_tokens", "fc_affine_video_1.weight",
                       "fc_affine_video_2.weight", "down_sampler.weight", "up_sampler.weight")

    def _packed_weights(self, cfg, params, xt, ot):
        """bf16 operand copies for the tensor-core path, re-packed only when a source parameter changed
        (torch bumps ``_version`` on every in-place update, e.g. optimizer.step / load_state_dict)."""
        if xt.dtype != torch.bfloat16:
            return None
        key = (tuple((params[n].data_ptr(), params[n]._version) for n in self._PACKED_SOURCES if n in params),
               tuple(xt.shape), tuple(ot.shape), str(xt.device))
        if getattr(self, "_packed_key", None) != key:
            self._packed_buf = adapter_forward_raw(cfg, params, xt, ot, pack_only=True)
            self._packed_key = key
        return self._packed_buf

    def forward(self, x, vis_token=None, caption=None, is_temporal=False):
        if caption is not None:
            raise NotImplementedError("the caption branch is dead code in the reference and is not implemented")
        if x.dim() != 4 or vis_token is None or vis_token.dim() != 4 or x.shape[-1] != 1 or vis_token.shape[-1] != 1:
            raise ValueError("expected x (BT,C,N,1) and vis_token (BT,C',N',1)")
        xt = x.squeeze(-1).permute(0, 2, 1)            # (BT,N,C): a view of the caller's contiguous buffer
        ot = vis_token.squeeze(-1).permute(0, 2, 1)
        # fp32 callers (the reference trains in fp32, main_trans.py:92) reach the tensor-core pipeline the way they would reach
        # torch's own bf16 kernels: under torch.autocast(dtype=bfloat16), or with module.compute_dtype = torch.bfloat16 -- the
        # tokens are cast at the boundary (differentiably) and the outputs come back in the caller's dtype.  Without either,
        # fp32 tensors take the exact-fp32 SIMT path.
        io_dtype = x.dtype
        low = getattr(self, "compute_dtype", None)
        if low is None and x.is_cuda and torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16:
            low = torch.bfloat16
        if low == torch.bfloat16 and xt.dtype == torch.float32:
            xt, ot = xt.to(torch.bfloat16), ot.to(torch.bfloat16)
        elif low not in (None, torch.bfloat16, torch.float32):
            raise TypeError(f"compute_dtype must be None, torch.float32 or torch.bfloat16 (got {low})")
        if xt.shape[1] != self.conv_dim_out or ot.shape[1] != self.conv_dim_in:
            raise RuntimeError(f"token counts ({xt.shape[1]}, {ot.shape[1]}) do not match conv_dim_out/conv_dim_in "
                               f"({self.conv_dim_out}, {self.conv_dim_in})")
        tensors = self._named_tensors()
        need_grad = torch.is_grad_enabled() and (x.requires_grad or vis_token.requires_grad or
                                                 any(p.requires_grad for _, p in tensors))
        if need_grad:
            named = [(n, p) for n, p in tensors if p.requires_grad and n in _GRAD_NAMES.values()]
            res = _AdapterFn.apply(self, tuple(n for n, _ in named), xt, ot, *[p for _, p in named])
        else:
            cfg, params = self._cfg(), self._param_dict()
            res = adapter_forward_raw(cfg, params, xt, ot, training=self.training,
                                      packed=self._packed_weights(cfg, params, xt, ot))
        out, smap = res[0], res[1]
        tmap = res[2] if len(res) > 2 else None
        if self.training and self.use_bn:
            self.bn1.num_batches_tracked += 1
            self.bn2.num_batches_tracked += 1
        if out.dtype != io_dtype:
            out, smap = out.to(io_dtype), smap.to(io_dtype)
        output = out.permute(0, 2, 1).unsqueeze(-1)
        spatial_att_maps = smap.unsqueeze(1)
        if VARIANT_DEFAULTS[self.variant]["temporal"]:
            T = VARIANT_DEFAULTS[self.variant]["T"]
            return output, spatial_att_maps, tmap.to(x.dtype).view(-1, T, 1, 1)
        return output, spatial_att_maps


class _PairFn(torch.autograd.Function):
    """Both adapters of one position through dgsct_adapter_pair_fwd / _pair_bwd_saved."""

    @staticmethod
    def forward(ctx, mod_a, mod_v, names_a, names_v, f_a, f_v, base_a, base_v, *flat):
        lib = L.lib()
        dev = f_a.device
        stream = torch.cuda.current_stream(dev).cuda_stream
        sides = []
        for mod, x, o in ((mod_a, f_a, f_v), (mod_v, f_v, f_a)):
            cfg, params = mod._cfg(), mod._param_dict()
            d, P, alive, x, o = _describe(cfg, params, x, o, mod.training)
            d.flags |= L.F_KEEP
            packed = mod._packed_weights(cfg, params, x, o)
            nbytes = lib.dgsct_workspace_bytes(C.byref(d))
            if nbytes == 0:
                raise ValueError("dgsct: invalid adapter description")
            ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
            sides.append(dict(cfg=cfg, params=params, d=d, P=P, alive=alive, x=x, o=o, packed=packed, ws=ws, nbytes=nbytes,
                              out=torch.empty_like(x), smap=torch.empty(x.shape[0], x.shape[1], dtype=x.dtype, device=dev)))
        A, V = sides
        pa = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)
        al = lambda t: C.c_void_p((t.data_ptr() + 255) // 256 * 256)
        bases = [b.contiguous() if b is not None else None for b in (base_a, base_v)]
        with torch.cuda.device(dev):
            rc = lib.dgsct_adapter_pair_fwd(C.byref(A["d"]), C.byref(A["P"]), pa(A["packed"]), C.byref(V["d"]), C.byref(V["P"]), pa(V["packed"]),
                                            pa(A["x"]), pa(V["x"]), pa(bases[0]), pa(bases[1]), pa(A["out"]), pa(V["out"]),
                                            pa(A["smap"]), pa(V["smap"]), None, None, al(A["ws"]), C.c_size_t(A["nbytes"]),
                                            al(V["ws"]), C.c_size_t(V["nbytes"]), C.c_void_p(stream))
        L.check(rc, "adapter_pair_fwd")
        ctx.mods, ctx.names, ctx.training = (mod_a, mod_v), (names_a, names_v), (mod_a.training, mod_v.training)
        ctx.packed, ctx.saved_ws = (A["packed"], V["packed"]), (A["ws"], V["ws"])
        ctx.has_base = (base_a is not None, base_v is not None)
        ctx.save_for_backward(A["x"], V["x"])
        return A["out"], V["out"], A["smap"], V["smap"]

    @staticmethod
    def backward(ctx, g_a, g_v, gs_a, gs_v):
        f_a, f_v = ctx.saved_tensors
        lib = L.lib()
        dev = f_a.device
        stream = torch.cuda.current_stream(dev).cuda_stream
        g_a = (g_a if g_a is not None else torch.zeros_like(f_a)).to(f_a.dtype).contiguous()
        g_v = (g_v if g_v is not None else torch.zeros_like(f_v)).to(f_v.dtype).contiguous()
        gs = [g.to(f_a.dtype).contiguous() if g is not None else None for g in (gs_a, gs_v)]
        sides, grads_out = [], []
        for mod, names, x, o, training in ((ctx.mods[0], ctx.names[0], f_a, f_v, ctx.training[0]),
                                           (ctx.mods[1], ctx.names[1], f_v, f_a, ctx.training[1])):
            cfg, params = mod._cfg(), mod._param_dict()
            d, P, alive, x, o = _describe(cfg, params, x, o, training)
            G = L.Grads()
            fresh, grads = [], {}
            fused = getattr(mod, "fused_grad_accumulation", False)
            tens = dict(mod._named_tensors())
            bicubic = cfg.get("resampler", "conv") == "bicubic"
            for field, name in _GRAD_NAMES.items():
                t = params.get(name)
                if t is None or (bicubic and field in ("resample_w", "resample_b")) or (field in ("t_w", "t_b") and not cfg.get("temporal", False)):
                    continue
                tgt = tens[name].grad if (fused and name in names and tens[name].is_leaf) else None
                if tgt is not None and tgt.dtype == torch.float32 and tgt.is_contiguous() and tgt.device == dev:
                    setattr(G, field, tgt.data_ptr())
                else:
                    fresh.append((field, name, t))
            if fresh:
                total = sum((t.numel() + 3) // 4 * 4 for _, _, t in fresh)
                flat = torch.zeros(total, dtype=torch.float32, device=dev)
                off = 0
                for field, name, t in fresh:
                    g = flat[off:off + t.numel()].view(t.shape)
                    grads[name] = g
                    setattr(G, field, g.data_ptr())
                    off += (t.numel() + 3) // 4 * 4
            nbytes = lib.dgsct_bwd_saved_workspace_bytes(C.byref(d))
            ws = _workspace(dev, ("pair_bwd", len(sides), stream), nbytes + 256)
            sides.append(dict(d=d, P=P, G=G, alive=alive, ws=ws, nbytes=nbytes, params=params))
            grads_out.append((names, grads, params))
        A, V = sides
        d_fa, d_fv, t_fa, t_fv = torch.empty_like(f_a), torch.empty_like(f_v), torch.empty_like(f_a), torch.empty_like(f_v)
        pa = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)
        al = lambda t: C.c_void_p((t.data_ptr() + 255) // 256 * 256)
        with torch.cuda.device(dev):
            rc = lib.dgsct_adapter_pair_bwd_saved(C.byref(A["d"]), C.byref(A["P"]), pa(ctx.packed[0]), al(ctx.saved_ws[0]),
                                                  C.byref(V["d"]), C.byref(V["P"]), pa(ctx.packed[1]), al(ctx.saved_ws[1]),
                                                  pa(f_a), pa(f_v), pa(g_a), pa(g_v), pa(gs[0]), pa(gs[1]), pa(d_fa), pa(d_fv),
                                                  pa(t_fa), pa(t_fv), C.byref(A["G"]), C.byref(V["G"]), al(A["ws"]),
                                                  C.c_size_t(A["nbytes"]), al(V["ws"]), C.c_size_t(V["nbytes"]), C.c_void_p(stream))
        L.check(rc, "adapter_pair_bwd")
        ctx.saved_ws = None
        flat_grads = []
        for names, grads, params in grads_out:
            for n in names:
                g, p_ = grads.get(n), params.get(n)
                flat_grads.append(g.view_as(p_) if (g is not None and p_ is not None) else None)
        for mod in ctx.mods:
            cb = getattr(mod, "_ddp_done_cb", None)
            if cb is not None:
                cb()
        return (None, None, None, None, d_fa, d_fv, g_a if ctx.has_base[0] else None, g_v if ctx.has_base[1] else None, *flat_grads)


def adapter_pair(audio_adapter: "VisualAdapter", vis_adapter: "VisualAdapter", f_a: torch.Tensor, f_v: torch.Tensor,
                 base_a: Optional[torch.Tensor] = None, base_v: Optional[torch.Tensor] = None):
    """Both adapters of one position in ONE engine call.  The interleave loop's

        f_a_res, _ = audio_adapter_blocks_pX[i](f_a.permute(0,2,1).unsqueeze(-1), f_v.permute(0,2,1).unsqueeze(-1))
        f_v_res, _ = vis_adapter_blocks_pX[i](f_v.permute(0,2,1).unsqueeze(-1), f_a.permute(0,2,1).unsqueeze(-1))
        ... f_v = <frozen block>(f_v) + f_v_res.squeeze(-1).permute(0,2,1) ; f_a likewise           (net_trans.py:891-906)

    becomes ``f_a, f_v, smap_a, smap_v = adapter_pair(aud, vis, f_a, f_v, base_a=<block output a>, base_v=<block output v>)``
    on token-major tensors f_a (BT,Na,Ca), f_v (BT,Nv,Cv): the two adapters run concurrently on forked streams and the
    residual add is folded into their last pass (``base_*`` None: the plain adapter outputs are returned).  Returns
    ``(out_a (BT,Na,Ca), out_v (BT,Nv,Cv), spatial_att_maps_a (BT,1,Na), spatial_att_maps_v (BT,1,Nv))``.
    Variants with a temporal gate (3-tuple adapters) are not covered by the pair entry."""
    for m in (audio_adapter, vis_adapter):
        if VARIANT_DEFAULTS[m.variant]["temporal"]:
            raise NotImplementedError("adapter_pair covers the 2-tuple variants (AVE / AVVP / AVQA / S4 / MS3)")
    if f_a.dim() != 3 or f_v.dim() != 3 or f_a.shape[0] != f_v.shape[0]:
        raise ValueError("expected token-major f_a (BT,Na,Ca) and f_v (BT,Nv,Cv)")
    ta, tv = audio_adapter._named_tensors(), vis_adapter._named_tensors()
    na = [(n, p) for n, p in ta if p.requires_grad and n in _GRAD_NAMES.values()]
    nv = [(n, p) for n, p in tv if p.requires_grad and n in _GRAD_NAMES.values()]
    out_a, out_v, sm_a, sm_v = _PairFn.apply(audio_adapter, vis_adapter, tuple(n for n, _ in na), tuple(n for n, _ in nv),
                                             f_a, f_v, base_a, base_v, *[p for _, p in na], *[p for _, p in nv])
    for m in (audio_adapter, vis_adapter):
        if m.training and m.use_bn:
            m.bn1.num_batches_tracked += 1
            m.bn2.num_batches_tracked += 1
    return out_a, out_v, sm_a.unsqueeze(1), sm_v.unsqueeze(1)


def patch(module, variant: str = "ave", compute_dtype=None):
    """Replace ``module.VisualAdapter`` (a reference task module such as ``nets.net_trans``) by the engine class
    bound to ``variant`` so the task script builds engine adapters without further changes.  ``compute_dtype=torch.bfloat16``
    sends an fp32 training script (the reference's default) through the bf16 tensor-core kernels."""
    import functools

    class _Bound(VisualAdapter):
        __init__ = functools.partialmethod(VisualAdapter.__init__, variant=variant)

    _Bound.compute_dtype = compute_dtype

    _Bound.__name__ = "VisualAdapter"
    module.VisualAdapter = _Bound
    return _Bound
