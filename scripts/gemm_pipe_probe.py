"""Which side of the tile pipeline costs ~0.9 us per k-block?  Runs the tcgen05 GEMM with the TMA loads and / or the MMAs
replaced by software mbarrier arrivals (no epilogue work in any mode) and sweeps the ring depth."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dgsct_b200 import _lib as L
lib = L.lib(); st = torch.cuda.current_stream().cuda_stream


def bench(M, N, K, batch, flags, stages=0, iters=20):
    A = torch.randn(batch, M, K).bfloat16().cuda()
    B = torch.randn(1, N, K).bfloat16().cuda()
    D = torch.empty(batch, M, N, device="cuda")
    def run():
        rc = lib.dgsct_debug_gemm_bf16(A.data_ptr(), B.data_ptr(), D.data_ptr(), M, N, K, batch, 0, 0, K, K, A.stride(0), 0, 1, flags, stages, st)
        assert rc == 0, rc
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): run()
    e1.record(); torch.cuda.synchronize()
    lib.dgsct_tc_error_flag(1)
    return e0.elapsed_time(e1) / iters * 1e3


SH = [("tok_logits aud s0", 4096, 32, 96, 80), ("fc aud s0", 4096, 96, 192, 80), ("chan vis s1", 576, 384, 384, 80),
      ("long K 4096x128x2304", 4096, 128, 2304, 20)]
BF = 1 << 16
MODES = [("full", BF), ("no epilogue", BF | (2 << 20)), ("no epi, no MMA", BF | ((2 | 4) << 20)), ("no epi, no TMA", BF | ((2 | 8) << 20)),
         ("barriers only", BF | ((2 | 4 | 8) << 20))]
for ctas in ("1", "2"):
    os.environ["DGSCT_TC_CTAS"] = ctas
    for nobres in ("", "1"):
        if nobres: os.environ["DGSCT_NO_BRES"] = "1"
        else: os.environ.pop("DGSCT_NO_BRES", None)
        print(f"==== CTAs/SM {ctas}  resident B {'off' if nobres else 'on'}")
        for name, M, N, K, batch in SH:
            tiles = batch * M // 128
            kblocks = tiles * ((K + 63) // 64)
            res = [f"{label}: {bench(M, N, K, batch, flags):7.1f}" for label, flags in MODES]
            sweep = [f"s{s}: {bench(M, N, K, batch, BF | (2 << 20), s):6.1f}" for s in (2, 3, 4, 6)]
            print(f"{name:22s} tiles {tiles:5d} k-blocks {kblocks:6d} | " + " | ".join(res) + " || no-epi stages " + " ".join(sweep), flush=True)
