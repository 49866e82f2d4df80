"""Per-shape, per-stage device time of one adapter BACKWARD call (incl. its forward recomputation), BT=80, bf16."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from dgsct_b200 import _lib as L
from dgsct_b200.adapter import adapter_forward_raw, adapter_backward_raw

dev = torch.device("cuda", 0)
gen = torch.Generator().manual_seed(0)
BT = 80
rows = {}
for kind in ("vis", "aud"):
    for s in range(4):
        C, N, Co, No = B.dims(kind, s)
        p = {k: v.to(dev) for k, v in B.make_params(C, N, Co, No, gen).items()}
        x = torch.randn(BT, N, C, generator=gen).bfloat16().to(dev)
        o = torch.randn(BT, No, Co, generator=gen).bfloat16().to(dev)
        do = (torch.randn(BT, N, C, generator=gen) * 1e-2).bfloat16().to(dev)
        dm = torch.randn(BT, N, generator=gen).bfloat16().to(dev)
        pk = adapter_forward_raw(B.ADAPTER_CFG, p, x, o, pack_only=True)
        for _ in range(2):
            adapter_backward_raw(B.ADAPTER_CFG, p, x, o, do, dm, None, training=True, packed=pk)
        torch.cuda.synchronize()
        L.profile_enable(True)
        R = 3
        for _ in range(R):
            adapter_backward_raw(B.ADAPTER_CFG, p, x, o, do, dm, None, training=True, packed=pk)
        torch.cuda.synchronize()
        prof = L.profile_collect()
        L.profile_enable(False)
        rows[f"{kind}_s{s}"] = {k: v[0] / R * 1e3 for k, v in prof.items()}
stages = sorted({k for r in rows.values() for k in r}, key=lambda k: -sum(r.get(k, 0) for r in rows.values()))
print(f"{'stage':16s}" + "".join(f"{n:>9s}" for n in rows) + "   (us per call)")
for st in stages:
    print(f"{st:16s}" + "".join(f"{rows[n].get(st, 0):9.1f}" for n in rows))
print(f"{'TOTAL':16s}" + "".join(f"{sum(rows[n].values()):9.1f}" for n in rows))
mult = {"s0": 4, "s1": 4, "s2": 12, "s3": 4}
print("weighted step total (ms):", sum(sum(rows[n].values()) * mult[n[-2:]] for n in rows) / 1e3)
