"""Which part of the adapter train step breaks CUDA-graph capture?  usage: capture_debug.py <mode> <what>
mode: global | thread_local | relaxed ; what: fwd | fwdbwd | fwdbwd_fused | full"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgsct_b200
from dgsct_b200.ddp import GradBucketer

mode, what = sys.argv[1], sys.argv[2]
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
opt = types.SimpleNamespace(is_multimodal=1, num_conv_group=2, is_before_layernorm=1, is_post_layernorm=1)
shapes = [(768, 64, 1536, 36), (1536, 36, 768, 64)] * 2
mods = []
for C, N, Co, No in shapes:
    m = dgsct_b200.VisualAdapter(C, C, "bottleneck", reduction_factor=8, opt=opt, num_tk=32, conv_dim_in=No, conv_dim_out=N,
                                 linear_in=Co, linear_out=C).to(dev).train()
    with torch.no_grad():
        m.gate.fill_(0.5); m.gate_av.fill_(0.5)
    m.gate_tk.requires_grad_(False)
    mods.append(m)
BT = 20
aud = torch.randn(BT, 64, 768, device=dev).bfloat16().requires_grad_(True)
vis = torch.randn(BT, 36, 1536, device=dev).bfloat16().requires_grad_(True)
bucketer = GradBucketer([[p for p in m.parameters() if p.requires_grad] for m in reversed(mods)])
adam = torch.optim.Adam([p for m in mods for p in m.parameters() if p.requires_grad], lr=1e-5, fused=True, capturable=True)
if what in ("fwdbwd_fused", "full"):
    for m in mods:
        m.fused_grad_accumulation = True


def step():
    if what == "fwd":
        with torch.no_grad():
            for i, m in enumerate(mods):
                x, o = (aud, vis) if i % 2 == 0 else (vis, aud)
                out, smap = m(x.permute(0, 2, 1).unsqueeze(-1), o.permute(0, 2, 1).unsqueeze(-1))
        return out.float().sum()
    bucketer.zero_grad()
    aud.grad = None; vis.grad = None
    outs = []
    for i, m in enumerate(mods):
        x, o = (aud, vis) if i % 2 == 0 else (vis, aud)
        outs.append(m(x.permute(0, 2, 1).unsqueeze(-1), o.permute(0, 2, 1).unsqueeze(-1)))
    for out, smap in reversed(outs):
        torch.autograd.backward([out, smap], [torch.ones_like(out) * 1e-2, torch.ones_like(smap)])
    outs.clear()
    if what == "full":
        bucketer.finish()
        adam.step()
    return aud.grad.float().abs().sum()


side = torch.cuda.Stream(dev)
side.wait_stream(torch.cuda.current_stream(dev))
with torch.cuda.stream(side):
    for _ in range(3):
        r = step()
torch.cuda.current_stream(dev).wait_stream(side)
torch.cuda.synchronize()
print(f"[{mode} {what}] eager ok, result {float(r):.4g}", flush=True)
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g, capture_error_mode=mode):
        r = step()
    g.replay(); g.replay()
    torch.cuda.synchronize()
    print(f"[{mode} {what}] CAPTURE OK, replay result {float(r):.4g}", flush=True)
except Exception as e:
    print(f"[{mode} {what}] CAPTURE FAILED: {type(e).__name__}: {str(e).splitlines()[0]}", flush=True)
