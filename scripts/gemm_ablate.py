import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgsct_b200 import _lib as L
lib = L.lib(); st = torch.cuda.current_stream().cuda_stream
def bench(M, N, K, batch, a_mn, b_mn, a_shared, b_shared, mode, flags, iters=10):
    na, nb = (1 if a_shared else batch), (1 if b_shared else batch)
    A = torch.randn(na, K, M).bfloat16().cuda() if a_mn else torch.randn(na, M, K).bfloat16().cuda()
    B = torch.randn(nb, K, N).bfloat16().cuda() if b_mn else torch.randn(nb, N, K).bfloat16().cuda()
    lda = M if a_mn else K; ldb = N if b_mn else K
    D = torch.empty(batch, M, N, device="cuda")
    sa = 0 if a_shared else A.stride(0); sb = 0 if b_shared else B.stride(0)
    def run():
        rc = lib.dgsct_debug_gemm_bf16(A.data_ptr(), B.data_ptr(), D.data_ptr(), M, N, K, batch, a_mn, b_mn, lda, ldb, sa, sb, mode, flags, 0, st)
        assert rc == 0, rc
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): run()
    e1.record(); torch.cuda.synchronize()
    lib.dgsct_tc_error_flag(1)
    return e0.elapsed_time(e1) / iters * 1e3
SH = [("conv aud s0", 4096, 192, 2304, 80, 0, 1, True, False), ("fc vis s0", 2304, 192, 96, 80, 0, 0, False, True), ("big", 4096, 256, 4096, 40, 0, 0, False, True)]
for name, *a in SH:
    res = []
    for label, flags in [("bf16 store", 1 << 16), ("fp32 store", 0), ("no store", (1 << 16) | (1 << 24)), ("store b=0", (1 << 16) | (1 << 25)),
                         ("no ld no store", (1 << 16) | (1 << 20) | (1 << 24)), ("no epilogue", (1 << 16) | (2 << 20))]:
        res.append(f"{label}: {bench(*a, 1, flags):7.1f}us")
    print(f"{name:12s} " + " | ".join(res), flush=True)
