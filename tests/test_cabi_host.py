"""CPU-side checks of the C ABI: the library loads, exports every symbol include/dgsct.h declares, and the
host-only entry points behave (no kernel is launched here)."""
import ctypes as C
import os
import re

import pytest
import torch

import dgsct_b200
from dgsct_b200 import _lib as L
from oracle import dgsct_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(L._SO):
        L.build()
    return L.lib()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "dgsct.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dgsct_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dgsct.h but not exported"
    assert declared == set(L.EXPORTS)
    assert lib.dgsct_version() == 100


def test_struct_layouts_match_header():
    hdr = open(os.path.join(ROOT, "include", "dgsct.h")).read()
    body = hdr[hdr.index("typedef struct dgsct_params {"):hdr.index("} dgsct_params;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.replace("const", "").replace("float", "").replace("int32_t", "").replace("typedef struct dgsct_params {", "")
        names += [n.strip().lstrip("*") for n in decl.split(",") if n.strip()]
    assert names == L.PARAM_FIELDS
    body = hdr[hdr.index("typedef struct dgsct_taps {"):hdr.index("} dgsct_taps;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    assert re.findall(r"float \*(\w+);", body) == L.TAP_FIELDS
    assert C.sizeof(L.Desc) == 17 * 4


def _desc(**kw):
    d = L.Desc()
    base = dict(BT=10, T=10, C=192, N=2304, Co=96, No=4096, K=32, r=8, g=2, io_dtype=L.DT_BF16, flags=0,
                alpha=0.3, beta=0.05, gamma=0.0, eps_ln=1e-5, eps_bn=1e-5, bn_momentum=0.1)
    base.update(kw)
    for k, v in base.items():
        setattr(d, k, v)
    return d


def test_workspace_bytes_and_validation(lib):
    n = lib.dgsct_workspace_bytes(C.byref(_desc()))
    assert n > 10 * 2304 * 192 * 2 * 3
    assert lib.dgsct_workspace_bytes(C.byref(_desc(C=191))) == 0           # odd width
    assert lib.dgsct_workspace_bytes(C.byref(_desc(r=7))) == 0             # C % r
    assert lib.dgsct_workspace_bytes(C.byref(_desc(io_dtype=5))) == 0
    assert lib.dgsct_workspace_bytes(C.byref(_desc(flags=L.F_TEMPORAL, BT=12))) == 0   # BT % T
    assert lib.dgsct_workspace_bytes(None) == 0
    # null pointers are reported, not dereferenced
    rc = lib.dgsct_adapter_fwd(C.byref(_desc()), None, None, None, None, None, None, None, None, None, 0, None)
    assert rc == -2 and lib.dgsct_status_string(rc) == b"null pointer"


@pytest.mark.parametrize("n_in,n_out", [(6, 4), (4, 6), (8, 8), (16, 12), (64, 48)])
def test_bicubic_csr_matches_torch_interpolate(lib, n_in, n_out):
    """The 16-tap stencil (CSR) and its transpose equal the dense matrix of torch's F.interpolate(mode='bicubic')."""
    ref = O.bicubic_matrix(n_in, n_out)
    cap = lib.dgsct_bicubic_csr_capacity(n_in, n_out)
    assert cap == 16 * n_out * n_out
    for tr, rows, cols in ((0, n_out * n_out, n_in * n_in), (1, n_in * n_in, n_out * n_out)):
        rp = torch.empty(rows + 1, dtype=torch.int32)
        ci = torch.empty(cap, dtype=torch.int32)
        va = torch.empty(cap, dtype=torch.float32)
        nnz = lib.dgsct_bicubic_csr(n_in, n_out, tr, C.c_void_p(rp.data_ptr()), C.c_void_p(ci.data_ptr()), C.c_void_p(va.data_ptr()))
        assert 0 < nnz <= cap and int(rp[0]) == 0 and int(rp[-1]) == nnz
        dense = torch.zeros(rows, cols, dtype=torch.float64)
        for r in range(rows):
            for j in range(int(rp[r]), int(rp[r + 1])):
                dense[r, int(ci[j])] += float(va[j])
        assert float((dense - (ref.t() if tr else ref)).abs().max()) < 2e-6
        if not tr:
            assert int((rp[1:] - rp[:-1]).max()) <= 16
    assert lib.dgsct_bicubic_csr(0, 4, 0, None, None, None) == -1


@pytest.mark.parametrize("n_in,n_out", [(6, 4), (4, 6), (8, 8), (16, 12), (12, 48)])
def test_bicubic_matrix_matches_torch_interpolate(lib, n_in, n_out):
    w = torch.empty(n_out * n_out, n_in * n_in, dtype=torch.float32)
    assert lib.dgsct_bicubic_matrix(n_in, n_out, C.c_void_p(w.data_ptr())) == 0
    ref = O.bicubic_matrix(n_in, n_out)
    assert float((w.double() - ref).abs().max()) < 2e-6
    assert torch.allclose(w.sum(1), torch.ones(n_out * n_out), atol=1e-5)


def test_module_surface_matches_reference_state_dict():
    import types
    from tests.golden_util import case_names, load_case
    variant_of = dict(ave_vis="ave", ave_aud="ave", ave_nobn="ave", avqa_vis="avqa", avvp_vis="avvp", s4_down="s4",
                      s4_up="s4", ms3_vis="ms3", pretrain_vis="pretrain", fewshot_aud="fewshot")
    for name in case_names():
        cfg, meta, params, t = load_case(name)
        opt = types.SimpleNamespace(is_multimodal=1, num_conv_group=cfg.g, is_before_layernorm=1, is_post_layernorm=1,
                                    num_tokens=cfg.K, alpha=cfg.alpha, beta=cfg.beta, gamma=cfg.gamma)
        ad = dgsct_b200.VisualAdapter(cfg.C, cfg.C, "bottleneck", reduction_factor=cfg.r, opt=opt, use_bn=cfg.use_bn,
                                      use_gate=cfg.use_gate, num_tk=cfg.K, conv_dim_in=cfg.No, conv_dim_out=cfg.N,
                                      linear_in=cfg.Co, linear_out=cfg.C, variant=variant_of[name])
        sd = ad.state_dict()
        ref_keys = set(params) | set(meta["extra_state_keys"])
        if variant_of[name] in ("s4", "ms3"):            # reference builds ln_before only if the option is set
            ref_keys |= {"ln_before.weight", "ln_before.bias"}
        assert set(sd) == ref_keys, (name, set(sd) ^ ref_keys)
        for k, v in params.items():
            assert tuple(sd[k].shape) == tuple(v.shape), (name, k)
        c = ad._cfg()
        for k in ("alpha", "beta", "gamma", "resampler", "ln_before", "ln_post", "gate_before_ln", "temporal", "K", "r", "g"):
            assert c[k] == getattr(cfg, k), (name, k, c[k], getattr(cfg, k))
        missing = ad.load_state_dict({k: v.float() for k, v in params.items()}, strict=False)
        assert not missing.unexpected_keys


def test_cpu_tensors_are_rejected_not_emulated():
    import types
    opt = types.SimpleNamespace(is_multimodal=1, num_conv_group=2, is_before_layernorm=1, is_post_layernorm=1)
    ad = dgsct_b200.VisualAdapter(32, 32, "bottleneck", reduction_factor=8, opt=opt, num_tk=4, conv_dim_in=25,
                                  conv_dim_out=16, linear_in=16, linear_out=32)
    x = torch.randn(2, 32, 16, 1)
    o = torch.randn(2, 16, 25, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        with torch.no_grad():
            ad(x, o)
    with pytest.raises(NotImplementedError):
        dgsct_b200.VisualAdapter(32, 32, "basic", reduction_factor=8, opt=opt)
    with pytest.raises(ValueError):
        with torch.no_grad():
            ad(x.squeeze(-1), o)


def test_keep_and_saved_workspace_sizes(lib):
    """DGSCT_F_KEEP selects the non-aliasing forward plan (larger); a backward from saved activations needs no recompute
    region (smaller than dgsct_bwd_workspace_bytes)."""
    d = _desc()
    plain = lib.dgsct_workspace_bytes(C.byref(d))
    d.flags = L.F_KEEP
    keep = lib.dgsct_workspace_bytes(C.byref(d))
    assert keep > plain > 0
    d.flags = 0
    full = lib.dgsct_bwd_workspace_bytes(C.byref(d))
    saved = lib.dgsct_bwd_saved_workspace_bytes(C.byref(d))
    assert 0 < saved < full
    assert full - saved >= keep - 4096          # the difference is the recompute region (the keep plan)
    bad = _desc(C=191)
    assert lib.dgsct_bwd_saved_workspace_bytes(C.byref(bad)) == 0


def test_fused_attention_launch_geometry_is_a_function_of_the_frame_shape():
    """Host logic of attn_fused.cuh (no device work): the forward's chunk rows and splits -- the quantities the bit-for-bit
    'a batch equals its clips' property rests on -- depend on (N, C, K) only; every plan fits the 227 KB of an SM; unsupported
    shapes are reported, not launched."""
    import ctypes as C
    lib = L.lib()

    def plan(kernel, BT, N, Cc, K=32):
        out = (C.c_int * 8)()
        assert lib.dgsct_debug_attn_plan(kernel, BT, N, Cc, K, out) == 0
        return list(out)

    shapes = [(192, 2304), (96, 4096), (384, 576), (192, 1024), (768, 144), (384, 256), (1536, 36), (768, 64), (320, 100), (1024, 49), (64, 10)]
    for Cc, N in shapes:
        ref = plan(0, 80, N, Cc)
        assert ref[0] == 1, (Cc, N)
        for BT in (1, 7, 10, 20, 80, 480):
            p = plan(0, BT, N, Cc)
            assert p[1:5] == ref[1:5], ("rows per chunk / buffers / chunks / splits must not depend on BT", Cc, N, BT, p, ref)
            assert p[1] % 16 == 0 and p[1] * p[3] >= N and p[4] <= p[3] and p[6] % 16 == 0 and p[5] * p[6] >= Cc
            for k in (0, 1, 2, 3):
                q = plan(k, BT, N, Cc)
                if q[0]:
                    assert 0 < q[7] <= 227 * 1024, (k, Cc, N, BT, q)
            assert plan(1, BT, N, Cc)[0] == 1 and plan(2, BT, N, Cc)[0] == 1
    # the token-softmax backward needs two 32 x C operand tiles next to a chunk: C = 1536 does not fit and is reported unsupported
    assert plan(3, 80, 36, 1536)[0] == 0 and plan(3, 80, 144, 768)[0] == 1
    # shapes outside the kernels' reach: channel count not a multiple of 32, more than 32 latent tokens
    assert plan(0, 8, 100, 100)[0] == 0 and plan(0, 8, 100, 96, 33)[0] == 0
    assert lib.dgsct_debug_attn_plan(9, 8, 100, 96, 32, (C.c_int * 8)()) < 0
