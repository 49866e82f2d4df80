"""Micro-benchmark of the tcgen05 GEMM on the adapter's dominant shapes (run under gpurun)."""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgsct_b200 import _lib as L

lib = L.lib()
st = torch.cuda.current_stream().cuda_stream


def bench(M, N, K, batch, a_mn, b_mn, a_shared, b_shared, mode, stages=0, out_bf16=1, iters=10):
    na, nb = (1 if a_shared else batch), (1 if b_shared else batch)
    A = torch.randn(na, K, M).bfloat16().cuda() if a_mn else torch.randn(na, M, K).bfloat16().cuda()
    B = torch.randn(nb, K, N).bfloat16().cuda() if b_mn else torch.randn(nb, N, K).bfloat16().cuda()
    lda = M if a_mn else K
    ldb = N if b_mn else K
    D = torch.empty(batch, M, N, device="cuda")
    sa = 0 if a_shared else A.stride(0)
    sb = 0 if b_shared else B.stride(0)
    def run():
        rc = lib.dgsct_debug_gemm_bf16(A.data_ptr(), B.data_ptr(), D.data_ptr(), M, N, K, batch, a_mn, b_mn, lda, ldb, sa, sb,
                                       mode, (out_bf16 << 16), stages, st)
        assert rc == 0, rc
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    assert lib.dgsct_tc_error_flag(1) == 0
    ms = e0.elapsed_time(e1) / iters
    return ms, 2.0 * M * N * K * batch / ms / 1e9


def torch_ref(M, N, K, batch, iters=10):
    A = torch.randn(M, K).bfloat16().cuda()
    B = torch.randn(batch, K, N).bfloat16().cuda()
    for _ in range(3):
        torch.matmul(A, B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        torch.matmul(A, B)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms, 2.0 * M * N * K * batch / ms / 1e9


SHAPES = [("conv aud s0", 4096, 192, 2304, 80, 0, 1, True, False), ("conv vis s0", 2304, 96, 4096, 80, 0, 1, True, False),
          ("conv vis s1", 576, 192, 1024, 80, 0, 1, True, False), ("fc vis s0", 2304, 192, 96, 80, 0, 0, False, True),
          ("fc aud s0", 4096, 96, 192, 80, 0, 0, False, True), ("chan vis s0", 2304, 192, 192, 80, 0, 0, False, True),
          ("big 4096x256x4096", 4096, 256, 4096, 40, 0, 0, False, True)]
for name, M, N, K, batch, a_mn, b_mn, ash, bsh in SHAPES:
    line = [f"{name:20s}"]
    for mode, stages in [(1, 0), (2, 110), (2, 72), (2, 54), (2, 40)]:
        try:
            ms, tf = bench(M, N, K, batch, a_mn, b_mn, ash, bsh, mode, stages)
            line.append(f"m{mode}/s{stages}: {ms*1e3:7.1f}us {tf:6.0f}TF")
        except AssertionError as e:
            line.append(f"m{mode}/s{stages}: ERR")
    if name.startswith("conv"):
        ms, tf = torch_ref(M, N, K, batch)
        line.append(f"cublas: {ms*1e3:7.1f}us {tf:6.0f}TF")
    print(" | ".join(line), flush=True)
