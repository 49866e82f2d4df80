"""ctypes binding of libdgsct.so (C ABI in include/dgsct.h).  No torch types cross this boundary."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "csrc", "libdgsct.so")

F_TRAIN, F_USE_BN, F_USE_GATE, F_LN_BEFORE, F_LN_POST, F_GATE_BEFORE_LN, F_TEMPORAL, F_NO_CONV_BIAS, F_FORCE_SIMT = (
    1 << 0, 1 << 1, 1 << 2, 1 << 3, 1 << 4, 1 << 5, 1 << 6, 1 << 7, 1 << 8)
F_KEEP = 1 << 11
DT_F32, DT_BF16 = 0, 1


class Desc(C.Structure):
    _fields_ = [("BT", C.c_int32), ("T", C.c_int32), ("C", C.c_int32), ("N", C.c_int32), ("Co", C.c_int32),
                ("No", C.c_int32), ("K", C.c_int32), ("r", C.c_int32), ("g", C.c_int32), ("io_dtype", C.c_int32),
                ("flags", C.c_uint32), ("alpha", C.c_float), ("beta", C.c_float), ("gamma", C.c_float),
                ("eps_ln", C.c_float), ("eps_bn", C.c_float), ("bn_momentum", C.c_float)]


PARAM_FIELDS = ["resample_w", "resample_b", "fc_w", "fc_b", "a1_w", "a1_b", "v1_w", "v1_b", "bott_w", "bott_b",
                "v2_w", "v2_b", "a2_w", "a2_b", "s_w", "s_b", "c_w", "c_b", "tokens", "gate_av", "gate",
                "down_w", "up_w", "bn1_w", "bn1_b", "bn2_w", "bn2_b", "bn1_rm", "bn1_rv", "bn2_rm", "bn2_rv",
                "lnb_w", "lnb_b", "lnp_w", "lnp_b", "t_w", "t_b",
                "rs_rowptr", "rs_col", "rs_val", "rst_rowptr", "rst_col", "rst_val"]


class Params(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in PARAM_FIELDS]


GRAD_FIELDS = ["resample_w", "resample_b", "fc_w", "fc_b", "a1_w", "a1_b", "v1_w", "v1_b", "bott_w", "bott_b", "v2_w", "v2_b",
               "a2_w", "a2_b", "s_w", "s_b", "c_w", "c_b", "tokens", "gate_av", "gate", "down_w", "up_w", "bn1_w", "bn1_b",
               "bn2_w", "bn2_b", "lnb_w", "lnb_b", "lnp_w", "lnp_b", "t_w", "t_b"]


class Grads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in GRAD_FIELDS]


TAP_FIELDS = ["V", "Tp", "X1", "obar", "m", "s", "Y", "Z"]


class Taps(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in TAP_FIELDS]


EXPORTS = ["dgsct_version", "dgsct_status_string", "dgsct_last_cuda_error", "dgsct_workspace_bytes",
           "dgsct_packed_bytes", "dgsct_pack_params", "dgsct_bwd_workspace_bytes", "dgsct_adapter_bwd",
           "dgsct_bwd_saved_workspace_bytes", "dgsct_adapter_bwd_saved",
           "dgsct_adapter_fwd", "dgsct_bicubic_matrix", "dgsct_num_stages", "dgsct_stage_name",
           "dgsct_profile_enable", "dgsct_profile_collect", "dgsct_launch_count", "dgsct_debug_gemm_bf16", "dgsct_debug_attn_plan",
           "dgsct_tc_error_flag", "dgsct_bicubic_csr_capacity", "dgsct_bicubic_csr",
           "dgsct_adam_tick", "dgsct_adam_flat", "dgsct_cast_bucket_bf16",
           "dgsct_adapter_pair_fwd", "dgsct_adapter_pair_bwd_saved",
           "dgsct_pool_tokens_fwd", "dgsct_pool_tokens_bwd", "dgsct_temporal_gates_fwd", "dgsct_temporal_gates_bwd"]

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libdgsct.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], capture_output=True, text=True)
    if verbose or out.returncode:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode:
        raise RuntimeError("building libdgsct.so failed")
    return _SO


def lib() -> C.CDLL:
    """Load the engine.  There is no fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise RuntimeError(f"dgsct_b200: {_SO} is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(or make -C dgsct_b200/csrc); there is no CPU/PyTorch fallback")
    L = C.CDLL(_SO)
    L.dgsct_version.restype = C.c_int
    L.dgsct_status_string.restype = C.c_char_p
    L.dgsct_status_string.argtypes = [C.c_int]
    L.dgsct_last_cuda_error.restype = C.c_char_p
    L.dgsct_workspace_bytes.restype = C.c_size_t
    L.dgsct_workspace_bytes.argtypes = [C.POINTER(Desc)]
    L.dgsct_adapter_fwd.restype = C.c_int
    L.dgsct_adapter_fwd.argtypes = [C.POINTER(Desc), C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.POINTER(Taps), C.c_void_p, C.c_size_t, C.c_void_p]
    L.dgsct_bwd_workspace_bytes.restype = C.c_size_t
    L.dgsct_bwd_workspace_bytes.argtypes = [C.POINTER(Desc)]
    L.dgsct_adapter_bwd.restype = C.c_int
    L.dgsct_adapter_bwd.argtypes = [C.POINTER(Desc), C.POINTER(Params)] + [C.c_void_p] * 8 + [C.POINTER(Grads), C.c_void_p,
                                                                                             C.c_size_t, C.c_void_p]
    L.dgsct_bwd_saved_workspace_bytes.restype = C.c_size_t
    L.dgsct_bwd_saved_workspace_bytes.argtypes = [C.POINTER(Desc)]
    L.dgsct_adapter_bwd_saved.restype = C.c_int
    L.dgsct_adapter_bwd_saved.argtypes = [C.POINTER(Desc), C.POINTER(Params)] + [C.c_void_p] * 9 + [C.POINTER(Grads), C.c_void_p,
                                                                                                   C.c_size_t, C.c_void_p]
    L.dgsct_packed_bytes.restype = C.c_size_t
    L.dgsct_packed_bytes.argtypes = [C.POINTER(Desc)]
    L.dgsct_pack_params.restype = C.c_int
    L.dgsct_pack_params.argtypes = [C.POINTER(Desc), C.POINTER(Params), C.c_void_p, C.c_void_p]
    L.dgsct_bicubic_matrix.restype = C.c_int
    L.dgsct_bicubic_matrix.argtypes = [C.c_int, C.c_int, C.c_void_p]
    L.dgsct_bicubic_csr_capacity.restype = C.c_int
    L.dgsct_bicubic_csr_capacity.argtypes = [C.c_int, C.c_int]
    L.dgsct_bicubic_csr.restype = C.c_int
    L.dgsct_bicubic_csr.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.dgsct_adapter_pair_fwd.restype = C.c_int
    L.dgsct_adapter_pair_fwd.argtypes = ([C.POINTER(Desc), C.POINTER(Params), C.c_void_p] * 2 + [C.c_void_p] * 10 +
                                         [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p])
    L.dgsct_adapter_pair_bwd_saved.restype = C.c_int
    L.dgsct_adapter_pair_bwd_saved.argtypes = ([C.POINTER(Desc), C.POINTER(Params), C.c_void_p, C.c_void_p] * 2 + [C.c_void_p] * 10 +
                                               [C.POINTER(Grads), C.POINTER(Grads), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                                C.c_void_p])
    L.dgsct_pool_tokens_fwd.restype = C.c_int
    L.dgsct_pool_tokens_fwd.argtypes = [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p]
    L.dgsct_pool_tokens_bwd.restype = C.c_int
    L.dgsct_pool_tokens_bwd.argtypes = [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_void_p]
    L.dgsct_temporal_gates_fwd.restype = C.c_int
    L.dgsct_temporal_gates_fwd.argtypes = [C.c_void_p] * 8 + [C.c_float] + [C.c_void_p] * 4 + [C.c_int] * 3 + [C.c_void_p]
    L.dgsct_temporal_gates_bwd.restype = C.c_int
    L.dgsct_temporal_gates_bwd.argtypes = [C.c_void_p] * 6 + [C.c_float] + [C.c_void_p] * 12 + [C.c_int] * 3 + [C.c_void_p]
    L.dgsct_adam_tick.restype = C.c_int
    L.dgsct_adam_tick.argtypes = [C.c_void_p, C.c_void_p]
    L.dgsct_adam_flat.restype = C.c_int
    L.dgsct_adam_flat.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t] + [C.c_float] * 6 + [C.c_void_p, C.c_void_p]
    L.dgsct_cast_bucket_bf16.restype = C.c_int
    L.dgsct_cast_bucket_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_void_p]
    L.dgsct_num_stages.restype = C.c_int
    L.dgsct_stage_name.restype = C.c_char_p
    L.dgsct_stage_name.argtypes = [C.c_int]
    L.dgsct_profile_enable.argtypes = [C.c_int]
    L.dgsct_profile_collect.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.c_int]
    L.dgsct_launch_count.restype = C.c_longlong
    L.dgsct_debug_attn_plan.restype = C.c_int
    L.dgsct_debug_attn_plan.argtypes = [C.c_int] * 5 + [C.POINTER(C.c_int)]
    L.dgsct_debug_gemm_bf16.restype = C.c_int
    L.dgsct_debug_gemm_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_longlong] * 4 + [C.c_int] * 3 + [C.c_void_p]
    L.dgsct_tc_error_flag.argtypes = [C.c_int]
    _lib = L
    return L


def profile_enable(on: bool) -> None:
    lib().dgsct_profile_enable(1 if on else 0)


def profile_collect() -> dict:
    """{stage: (ms, calls)} accumulated since the last collect; synchronises on the recorded events."""
    L = lib()
    n = L.dgsct_num_stages()
    ms = (C.c_double * n)()
    calls = (C.c_longlong * n)()
    check(L.dgsct_profile_collect(ms, calls, n), "profile_collect")
    return {L.dgsct_stage_name(i).decode(): (ms[i], calls[i]) for i in range(n) if calls[i]}


def launch_count() -> int:
    return int(lib().dgsct_launch_count())


def check(status: int, what: str) -> None:
    if status == 0:
        return
    L = lib()
    msg = L.dgsct_status_string(status).decode()
    if status == -4:
        msg += ": " + L.dgsct_last_cuda_error().decode()
    if status in (-1,):
        raise ValueError(f"dgsct {what}: {msg}")
    raise RuntimeError(f"dgsct {what}: {msg}")
