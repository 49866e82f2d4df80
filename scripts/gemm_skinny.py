"""Ablation of the tcgen05 GEMM on the adapter's short-K (HBM-bound) products: which part of the tile pipeline
limits it (loads / epilogue TMEM reads / global stores) and whether 2 CTAs per SM help.  Run under gpurun."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgsct_b200 import _lib as L
lib = L.lib(); st = torch.cuda.current_stream().cuda_stream


def bench(M, N, K, batch, flags, iters=20):
    A = torch.randn(batch, M, K).bfloat16().cuda()
    B = torch.randn(1, N, K).bfloat16().cuda()
    D = torch.empty(batch, M, N, device="cuda")
    def run():
        rc = lib.dgsct_debug_gemm_bf16(A.data_ptr(), B.data_ptr(), D.data_ptr(), M, N, K, batch, 0, 0, K, K, A.stride(0), 0, 1, flags, 0, st)
        assert rc == 0, rc
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): run()
    e1.record(); torch.cuda.synchronize()
    assert lib.dgsct_tc_error_flag(1) == 0
    return e0.elapsed_time(e1) / iters * 1e3


SH = [("tok_logits aud s0", 4096, 32, 96, 80), ("chan aud s0", 4096, 96, 96, 80), ("fc aud s0", 4096, 96, 192, 80),
      ("chan vis s0", 2304, 192, 192, 80), ("chan vis s1", 576, 384, 384, 80), ("chan vis s2", 144, 768, 768, 80),
      ("chan vis s3", 36, 1536, 1536, 80), ("chan aud s2", 256, 384, 384, 80)]
MODES = [("bf16 store", 1 << 16), ("fp32 store", 0), ("no store", (1 << 16) | (1 << 24)),
         ("no ld no store", (1 << 16) | (1 << 20) | (1 << 24)), ("no epilogue", (1 << 16) | (2 << 20))]
for env in sys.argv[1:] or ["", "DGSCT_TC_CTAS=2"]:
    for kv in env.split(","):
        if "=" in kv:
            k, v = kv.split("="); os.environ[k] = v
    print(f"==== env [{env}]")
    for name, M, N, K, batch in SH:
        rd = batch * M * K * 2 / 1e6
        res = [f"{label}: {bench(M, N, K, batch, flags):7.1f}us" for label, flags in MODES]
        print(f"{name:18s} A={rd:6.1f}MB ({rd / 6.5e3 * 1e3:5.1f}us@6.5TB/s) " + " | ".join(res), flush=True)
    for kv in env.split(","):
        if "=" in kv:
            os.environ.pop(kv.split("=")[0], None)
