"""One adapter call (forward with kept activations + backward) at a reference shape, bracketed by cudaProfilerStart/Stop,
for `ncu --profile-from-start off` captures.  usage: python scripts/one_call.py <shape> [BT] [fwd|train]"""
import dataclasses
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dgsct_b200.adapter import adapter_backward_raw, adapter_forward_raw      # noqa: E402

SHAPES = {"vis_s0": (192, 2304, 96, 4096), "vis_s1": (384, 576, 192, 1024), "vis_s2": (768, 144, 384, 256),
          "vis_s3": (1536, 36, 768, 64), "aud_s0": (96, 4096, 192, 2304), "aud_s1": (192, 1024, 384, 576),
          "aud_s2": (384, 256, 768, 144), "aud_s3": (768, 64, 1536, 36)}


def main():
    sys.path.insert(0, ROOT)
    import bench as B
    shape = sys.argv[1] if len(sys.argv) > 1 else "aud_s2"
    BT = int(sys.argv[2]) if len(sys.argv) > 2 else 80
    mode = sys.argv[3] if len(sys.argv) > 3 else "train"
    C, N, Co, No = SHAPES[shape]
    gen = torch.Generator().manual_seed(1)
    p = {k: v.cuda() for k, v in B.make_params(C, N, Co, No, gen).items()}
    x = torch.randn(BT, N, C, generator=gen).bfloat16().cuda()
    o = torch.randn(BT, No, Co, generator=gen).bfloat16().cuda()
    d_out = (torch.randn(BT, N, C, generator=gen) * 1e-2).bfloat16().cuda()
    d_smap = torch.randn(BT, N, generator=gen).bfloat16().cuda()
    cfg = B.ADAPTER_CFG
    pk = adapter_forward_raw(cfg, p, x, o, pack_only=True)
    grads = {k: torch.zeros_like(v) for k, v in p.items() if "running" not in k}
    os.environ["DGSCT_NO_SIDE"] = "1"

    def call():
        if mode == "fwd":
            adapter_forward_raw(cfg, p, x, o, training=False, packed=pk)
            return
        keep = []
        adapter_forward_raw(cfg, p, x, o, training=True, packed=pk, keep=keep)
        adapter_backward_raw(cfg, p, x, o, d_out, d_smap, None, training=True, packed=pk, saved=keep[0], grad_targets=grads)

    for _ in range(2):
        call()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    call()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
