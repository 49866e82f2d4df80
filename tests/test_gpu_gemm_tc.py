"""Bring-up tests of the tcgen05/TMA GEMM (gemm_tc.cuh) against torch fp32 matmul of the same bf16 operands."""
import ctypes as C

import pytest
import torch

from dgsct_b200 import _lib as L

pytestmark = pytest.mark.gpu


def run_gemm(M, N, K, batch, a_mn, b_mn, a_shared=False, b_shared=False, use_tc=1, bn_hint=0, stages=0, seed=0):
    g = torch.Generator().manual_seed(seed)
    pad = lambda x: (x + 7) // 8 * 8
    na, nb = (1 if a_shared else batch), (1 if b_shared else batch)
    if a_mn:
        lda = pad(M); A = torch.zeros(na, K, lda); A[:, :, :M] = torch.randn(na, K, M, generator=g)
        Aref = A[:, :, :M].transpose(1, 2)
    else:
        lda = pad(K); A = torch.zeros(na, M, lda); A[:, :, :K] = torch.randn(na, M, K, generator=g)
        Aref = A[:, :, :K]
    if b_mn:
        ldb = pad(N); B = torch.zeros(nb, K, ldb); B[:, :, :N] = torch.randn(nb, K, N, generator=g)
        Bref = B[:, :, :N].transpose(1, 2)
    else:
        ldb = pad(K); B = torch.zeros(nb, N, ldb); B[:, :, :K] = torch.randn(nb, N, K, generator=g)
        Bref = B[:, :, :K]
    A16, B16 = A.bfloat16().cuda(), B.bfloat16().cuda()
    D = torch.full((batch, M, N), float("nan"), device="cuda")
    sa = 0 if a_shared else A16.stride(0)
    sb = 0 if b_shared else B16.stride(0)
    lib = L.lib()
    rc = lib.dgsct_debug_gemm_bf16(A16.data_ptr(), B16.data_ptr(), D.data_ptr(), M, N, K, batch, a_mn, b_mn, lda, ldb, sa, sb,
                                   use_tc, bn_hint, stages, torch.cuda.current_stream().cuda_stream)
    L.check(rc, "debug_gemm")
    torch.cuda.synchronize()
    assert lib.dgsct_tc_error_flag(1) == 0, "a pipeline wait timed out"
    ref = torch.matmul(Aref.bfloat16().float().expand(batch, -1, -1), Bref.bfloat16().float().expand(batch, -1, -1).transpose(1, 2))
    err = float((D.cpu() - ref).abs().max() / ref.abs().max())
    return err


CASES = [
    # M, N, K, batch, a_mn, b_mn, a_shared, b_shared
    (128, 64, 64, 1, 0, 0, False, False),
    (256, 256, 512, 2, 0, 0, False, True),
    (200, 96, 100, 3, 0, 0, False, True),
    (36, 192, 768, 3, 0, 0, False, True),        # ragged M (stage-3 tokens), linear weights
    (144, 32, 768, 2, 0, 0, False, True),        # token logits: N = 32 latent tokens
    (300, 384, 96, 2, 0, 0, False, True),        # N > 256 -> two 192-wide tiles
    (256, 128, 512, 2, 0, 1, True, False),       # resampler: shared K-major weights x per-frame MN-major tokens
    (256, 96, 520, 2, 0, 1, True, False),        # N = 96 -> 64B-swizzled MN-major tile
    (256, 192, 256, 2, 0, 1, True, False),
    (144, 256, 256, 2, 0, 1, True, False),
    (64, 48, 36, 2, 0, 1, True, False),          # stage-3 audio resampler: K = 36 tokens
    (300, 96, 200, 3, 0, 1, True, False),        # frame packing (2 frames per tile) with an odd batch
    (130, 48, 70, 7, 0, 1, True, False),         # frame packing, 4 frames per tile, ragged last group
    (260, 128, 130, 5, 0, 1, True, False),       # frame packing of 128-wide frames
    (192, 32, 300, 2, 1, 1, False, False),       # latent-token update: both operands token-major
    (100, 192, 32, 2, 0, 1, False, False),       # x update: K = 32
    (1536, 32, 36, 2, 1, 1, False, False),
    # short-K, shared B: resident B + staged (row-contiguous) stores + frames flattened into one M extent
    (36, 40, 96, 7, 0, 0, False, True),          # 36-token frames: 32-row slabs straddle frames
    (70, 90, 40, 5, 0, 0, False, True),          # output rows not 16-byte aligned (scalar tail of the staged store)
    (144, 200, 130, 3, 0, 0, False, True),       # ragged N and K
    (4096, 32, 96, 3, 0, 0, False, True),        # token logits at the stage-0 audio shape
    (576, 384, 384, 2, 0, 0, False, True),       # two n-blocks, B too large to stay resident next to a second CTA
    (64, 1536, 192, 4, 0, 0, False, True),       # six n-blocks, resident B covers all of them
    (300, 96, 64, 3, 0, 0, False, False),        # per-frame B: no flattening, no resident B
]


@pytest.mark.parametrize("case", CASES, ids=[str(c) for c in CASES])
def test_tc_gemm_matches_fp32_matmul(case):
    M, N, K, batch, a_mn, b_mn, ash, bsh = case
    err = run_gemm(M, N, K, batch, a_mn, b_mn, ash, bsh, use_tc=1)
    assert err < 2e-3, err        # bf16 products are exact in fp32; only accumulation order differs


def test_simt_debug_gemm_agrees():
    assert run_gemm(200, 96, 100, 3, 0, 1, True, False, use_tc=0) < 1e-5
