#!/usr/bin/env python
"""bench.py -- AVE clips/sec of the DG-SCT adapter hot path (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic clips: the 48 ``VisualAdapter.forward`` calls
of one AVE forward (12 paired layers x {audio,visual} x {p1,p2}, DG-SCT/AVE/nets/net_trans.py:878-906) at the
reference geometry (Swin-V2-L/192 + HTS-AT/256 token shapes), B clips x T frames per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W]            engine arm (CUDA, C ABI)
  python bench.py --impl reference [...]                         reference arm: the CPU oracle on host cores

Under torchrun (N>1) every rank processes its own B clips (weak scaling, no data-path collective in the
forward); the timed region is bracketed by barrier + synchronize and the max over ranks is reported.
"""
import argparse
import dataclasses
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PAIRS = [0, 0, 1, 1, 2, 2, 2, 2, 2, 2, 3, 3]                       # stage of each paired layer (net_trans.py:882-885)
VIS = [(192, 2304), (384, 576), (768, 144), (1536, 36)]            # (C, N) per Swin-V2-L/192 stage
AUD = [(96, 4096), (192, 1024), (384, 256), (768, 64)]             # (C, N) per HTS-AT/256 stage
ADAPTER_CFG = dict(K=32, r=8, g=2, T=10, alpha=0.3, beta=0.05, gamma=0.0, resampler="conv", ln_before=True,
                   ln_post=True, gate_before_ln=False, use_bn=True, use_gate=True, temporal=False)


# Workloads = BASELINE.json configs: "ave" is the metric's own (configs 1-3, Swin-V2-L/192 token shapes: the geometry the
# reference really runs); "b224" is BASELINE.json's Swin-V2-B/224 wording; "s4" is config 4 (AVS-S4: bicubic
# resampler, no ln_before, gate before ln_post, BN / gate off, B=4 clips x 5 frames, PVT_AVSModel.py:90-316, avs_s4/train.sh);
# "avqa_t60" is config 5 (AVQA copy: K=2 latent tokens, g=4, BN / gate off, net_avst.py:27-218, T=60 frames per clip).
VIS_B224 = [(128, 3136), (256, 784), (512, 196), (1024, 49)]
WORKLOADS = {
    "ave": dict(variant="ave", vis=VIS, B=8, T=10, cfg=ADAPTER_CFG, what="AVE", module_kw=dict(use_bn=True, use_gate=True, num_tk=32),
                opt=dict(num_conv_group=2)),
    "b224": dict(variant="ave", vis=VIS_B224, B=8, T=10, cfg=ADAPTER_CFG, what="AVE, Swin-V2-B/224 token shapes",
                 module_kw=dict(use_bn=True, use_gate=True, num_tk=32), opt=dict(num_conv_group=2)),
    "s4": dict(variant="s4", vis=VIS, B=4, T=5, what="AVS-S4",
               cfg=dict(ADAPTER_CFG, T=5, resampler="bicubic", ln_before=False, gate_before_ln=True, use_bn=False, use_gate=False),
               module_kw=dict(use_bn=False, use_gate=False), opt=dict(num_conv_group=2, num_tokens=32)),
    "avqa_t60": dict(variant="avqa", vis=VIS, B=8, T=60, what="AVQA long-temporal",
                     cfg=dict(ADAPTER_CFG, T=60, K=2, g=4, use_bn=False, use_gate=False),
                     module_kw=dict(use_bn=False, use_gate=False), opt=dict(num_conv_group=4, num_tokens=2)),
}
_WL = {"name": "ave"}


def wl():
    return WORKLOADS[_WL["name"]]


def metric_name(args):
    w = wl()
    what = ("adapter-stack train step (fwd+bwd+all-reduce+Adam over the 48 DG-SCT adapters)" if args.mode == "train"
            else "adapter-stack forward (48 VisualAdapter.forward calls)")
    return f"{w['what']} clips/sec (B={args.B},T={args.T}): {what}"


def layer_list():
    out = []
    for li, s in enumerate(PAIRS):
        for pos in ("p1", "p2"):
            out.append(("aud", li, pos, s))
            out.append(("vis", li, pos, s))
    return out


def dims(kind, s):
    vis = wl()["vis"]
    (C, N), (Co, No) = (vis[s], AUD[s]) if kind == "vis" else (AUD[s], vis[s])
    return C, N, Co, No


def alg_bytes(C, N, Co, No, BT, K=32, r=8, g=2, s_act=2, s_w=2):
    P = (N * No + N + Co * C + C + 2 * (C * C + C) + 3 * (C * C // 2 + C // 2) + C // 2 + 1 + C * C // 2 + C
         + K * C + 3 + 2 * C * (C // r) // g + 2 * (C // r) + 6 * C)
    return s_act * BT * (2 * N * C + No * Co + N) + s_w * P


def alg_flops(C, N, Co, No, BT, K=32, r=8, g=2):
    return (2 * BT * (N * No * Co + N * Co * C) + 8 * BT * K * N * C + 2 * BT * N * C * C + BT * N * C * C
            + 4 * BT * N * C * (C // r) // g)


def init_param(shape, fan_in, gen, kind):
    if kind == "gate":
        return torch.rand(shape, generator=gen) * 0.75 + 0.25
    if kind == "tokens":
        return torch.rand(shape, generator=gen)
    if kind == "var":
        return torch.rand(shape, generator=gen) + 0.5
    if kind == "mean":
        return torch.randn(shape, generator=gen) * 0.1
    if kind == "one":
        return 1.0 + 0.1 * torch.randn(shape, generator=gen)
    if kind == "zero":
        return 0.1 * torch.randn(shape, generator=gen)
    b = 1.0 / max(fan_in, 1) ** 0.5
    return (torch.rand(shape, generator=gen) * 2 - 1) * b


def make_params(C, N, Co, No, gen, K=None, r=8, g=None):
    """Reference initialisers with gates / BN buffers randomised (SURVEY.md §8d 'Synthetic inputs')."""
    K = wl()["cfg"]["K"] if K is None else K
    g = wl()["cfg"]["g"] if g is None else g
    H, Cd = C // 2, C // r
    spec = {
        "conv_adapter.weight": ((N, No, 1, 1), No, "w"), "conv_adapter.bias": ((N,), No, "w"),
        "fc.weight": ((C, Co), Co, "w"), "fc.bias": ((C,), Co, "w"),
        "fc_affine_audio_1.weight": ((C, C), C, "w"), "fc_affine_audio_1.bias": ((C,), C, "w"),
        "fc_affine_video_1.weight": ((C, C), C, "w"), "fc_affine_video_1.bias": ((C,), C, "w"),
        "fc_affine_bottleneck.weight": ((H, C), C, "w"), "fc_affine_bottleneck.bias": ((H,), C, "w"),
        "fc_affine_video_2.weight": ((H, C), C, "w"), "fc_affine_video_2.bias": ((H,), C, "w"),
        "fc_affine_audio_2.weight": ((H, C), C, "w"), "fc_affine_audio_2.bias": ((H,), C, "w"),
        "fc_affine_v_s_att.weight": ((1, H), H, "w"), "fc_affine_v_s_att.bias": ((1,), H, "w"),
        "fc_affine_v_c_att.weight": ((C, H), H, "w"), "fc_affine_v_c_att.bias": ((C,), H, "w"),
        "my_tokens": ((K, C), 0, "tokens"), "gate_av": ((1,), 0, "gate"), "gate": ((1,), 0, "gate"),
        "down_sampler.weight": ((Cd, C // g, 1, 1), C // g, "w"), "up_sampler.weight": ((C, Cd // g, 1, 1), Cd // g, "w"),
        "bn1.weight": ((Cd,), 0, "one"), "bn1.bias": ((Cd,), 0, "zero"), "bn2.weight": ((C,), 0, "one"), "bn2.bias": ((C,), 0, "zero"),
        "bn1.running_mean": ((Cd,), 0, "mean"), "bn1.running_var": ((Cd,), 0, "var"),
        "bn2.running_mean": ((C,), 0, "mean"), "bn2.running_var": ((C,), 0, "var"),
        "ln_before.weight": ((C,), 0, "one"), "ln_before.bias": ((C,), 0, "zero"),
        "ln_post.weight": ((C,), 0, "one"), "ln_post.bias": ((C,), 0, "zero"),
    }
    return {k: init_param(shp, fan, gen, kind) for k, (shp, fan, kind) in spec.items()}


def make_tokens(B, T, gen, dtype):
    BT = B * T
    vis = [torch.randn(BT, N, C, generator=gen).to(dtype) for (C, N) in wl()["vis"]]
    aud = [torch.randn(BT, N, C, generator=gen).to(dtype) for (C, N) in AUD]
    return vis, aud


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            txt = self.p.communicate(timeout=5)[0]
        except Exception:
            txt = ""
        sm, mx, reasons = [], [], set()
        for line in txt.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def oracle_cfg(O, C, N, Co, No):
    """Oracle AdapterConfig of the current workload (variant presets + the workload's overrides)."""
    c = wl()["cfg"]
    return O.make_config(wl()["variant"], C, N, Co, No, K=c["K"], r=c["r"], g=c["g"], T=c["T"], use_bn=c["use_bn"], use_gate=c["use_gate"])


def make_cpu_step(args, B, device="cpu", autocast=False):
    """One step of the same workload on B clips with the ORACLE (the restatement of the reference path pinned to the
    reference's own classes by tests/golden): in train mode forward + backward (torch autograd of the oracle) + Adam, with
    train-mode BatchNorm.  device='cpu' is the reference arm / cpu_baseline; device='cuda' (+ bf16 autocast) is the
    PyTorch-eager-on-B200 comparator.  Baselines only: the engine path never calls this."""
    from oracle import dgsct_oracle as O
    gen = torch.Generator().manual_seed(43)
    vis, aud = make_tokens(B, args.T, gen, torch.float32)
    vis = [t.to(device) for t in vis]
    aud = [t.to(device) for t in aud]
    stack = []
    for kind, li, pos, s in layer_list():
        C, N, Co, No = dims(kind, s)
        cfg = oracle_cfg(O, C, N, Co, No)
        names = set(O.param_shapes(cfg))
        stack.append((kind, s, cfg, {k: v.to(device) for k, v in make_params(C, N, Co, No, gen).items() if k in names}))
    ac = (lambda: torch.autocast(device, dtype=torch.bfloat16)) if autocast else (lambda: torch.autocast(device, enabled=False))
    if args.mode != "train":
        def step():
            with torch.no_grad(), ac():
                for kind, s, cfg, p in stack:
                    x, o = (vis[s], aud[s]) if kind == "vis" else (aud[s], vis[s])
                    O.adapter_forward(cfg, p, x, o, training=False)
        return step
    leaves = []
    for kind, s, cfg, p in stack:
        for k, v in p.items():
            if "running" not in k:
                v.requires_grad_(True)
                leaves.append(v)
    opt = torch.optim.Adam(leaves, lr=1e-5)
    d_vis = [(torch.randn(t.shape, generator=gen) * 1e-2).to(device) for t in vis]
    d_aud = [(torch.randn(t.shape, generator=gen) * 1e-2).to(device) for t in aud]

    def step():
        opt.zero_grad(set_to_none=True)
        for kind, s, cfg, p in stack:
            x, o = (vis[s], aud[s]) if kind == "vis" else (aud[s], vis[s])
            with ac():
                out, smap, _, stats = O.adapter_forward(cfg, p, x, o, training=True)
            ((out.float() * (d_vis[s] if kind == "vis" else d_aud[s])).sum() + smap.float().sum()).backward()
            for k, v in stats.items():
                p[k] = v
        opt.step()
    return step


REF_CLIPS = {"ave": None, "b224": None, "s4": None, "avqa_t60": 1}      # None: the engine's own B (same config)


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (the oracle restatement pinned to the reference's classes by
    tests/golden), all host threads, on the engine arm's config: B clips x T frames per step (the AVQA T=60 workload is
    bounded to 1 clip = 60 frames per step and says so)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = REF_CLIPS[_WL["name"]] or args.B
    if args.ref_clips:
        Bs = args.ref_clips
    step = make_cpu_step(args, Bs)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    v = Bs / dt
    sample = (f"B={Bs} clips (BT={Bs * args.T}) per step" + ("" if Bs == args.B else f" -- bounded sample of the B={args.B} config")
              + f", 48 adapters, fp32, {args.mode}, torch CPU ({cores} threads)")
    line = {"metric": metric_name(args), "value": v, "unit": "clips/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference",
            "config": workload_config(args, "cpu"),
            "cpu_baseline": {"value": v, "unit": "clips/s", "cores": cores, "kind": "port", "clips_per_step": Bs, "sample": sample},
            "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, where, launch=None):
    what = ("train step (forward + backward + grad all-reduce + fused Adam)" + (", train-mode BatchNorm" if wl()["cfg"]["use_bn"] else "")
            if args.mode == "train" else "forward, eval mode")
    c = wl()["cfg"]
    geom = "Swin-V2-B/224" if _WL["name"] == "b224" else "Swin-V2-L/192"
    return {"workload": f"{_WL['name']}: {wl()['what']} adapter stack {what}: 48 VisualAdapter calls, {geom} + HTS-AT/256 token shapes, "
                        f"B={args.B} clips x T={args.T} frames per GPU, K={c['K']} r={c['r']} g={c['g']}, resampler={c['resampler']}",
            "B_per_gpu": args.B, "T": args.T, "adapters": 48, "mode": args.mode,
            "l2": "per-step working set (packed bf16 weights + tokens + intermediates, > 600 MB) exceeds the 126 MB L2; no explicit flush",
            "launch": launch or "host loop",
            "device": where}


# ------------------------------------------------------------------------------------------------ engine arm
def run_engine(args, rank, world, local_rank):
    import dgsct_b200
    from dgsct_b200 import _lib as L
    from dgsct_b200.adapter import adapter_forward_raw

    assert torch.cuda.is_available(), "the engine arm needs a B200; there is no CPU fallback"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    CFG = wl()["cfg"]
    # CPU baseline first, on rank 0, BEFORE the process group exists: the other ranks then wait inside the rendezvous
    # (blocking socket reads) instead of spinning in an NCCL barrier on the same host cores (round 1: 0.012 clips/s at N=8).
    cpu_base = cpu_baseline(args) if (rank == 0 and not args.no_cpu_baseline) else None
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist_.init_process_group("nccl", device_id=dev)
        dist = dist_
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    gen = torch.Generator().manual_seed(43 + rank)
    vis_h, aud_h = make_tokens(args.B, args.T, gen, dtype)
    vis_h = [t.pin_memory() for t in vis_h]
    aud_h = [t.pin_memory() for t in aud_h]
    vis = [t.to(dev) for t in vis_h]
    aud = [t.to(dev) for t in aud_h]
    pgen = torch.Generator().manual_seed(1234)          # identical (replicated) parameters on every rank
    stack = []
    import types
    opt = types.SimpleNamespace(is_multimodal=1, is_before_layernorm=1, is_post_layernorm=1, **wl()["opt"])
    modules = []
    for kind, li, pos, s in layer_list():
        C, N, Co, No = dims(kind, s)
        p = {k: v.to(dev) for k, v in make_params(C, N, Co, No, pgen).items()}
        x0, o0 = (vis[s], aud[s]) if kind == "vis" else (aud[s], vis[s])
        pk = adapter_forward_raw(CFG, p, x0, o0, pack_only=True) if dtype == torch.bfloat16 else None
        stack.append((kind, s, p, pk))
        m = dgsct_b200.VisualAdapter(C, C, "bottleneck", reduction_factor=8, opt=opt, conv_dim_in=No, conv_dim_out=N,
                                     linear_in=Co, linear_out=C, variant=wl()["variant"], **wl()["module_kw"])
        m.load_state_dict(p, strict=False)
        modules.append(m.to(dev).eval())
    BT = args.B * args.T

    # The audio and visual adapters of one position read the same two token tensors and are independent
    # (net_trans.py:891-892): they are issued on two streams and joined after every pair, as the interleave loop's
    # data dependencies allow (the frozen backbone blocks sit between consecutive pairs).
    s_aud, s_vis = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def step_resident():
        cur = torch.cuda.current_stream(dev)
        for i in range(0, len(stack), 2):
            s_aud.wait_stream(cur)
            s_vis.wait_stream(cur)
            for (kind, s, p, pk), strm in ((stack[i], s_aud), (stack[i + 1], s_vis)):
                x, o = (vis[s], aud[s]) if kind == "vis" else (aud[s], vis[s])
                with torch.cuda.stream(strm):
                    adapter_forward_raw(CFG, p, x, o, training=False, packed=pk)
            cur.wait_stream(s_aud)
            cur.wait_stream(s_vis)

    h2d = sum(t.numel() * t.element_size() for t in vis_h + aud_h)
    d2h_buf = torch.empty(2 * BT * 64 + 1, dtype=torch.float32).pin_memory()
    # static device inputs of the end-to-end path: every e2e step copies the pinned host tokens into them
    vis_in = [torch.empty_like(t) for t in vis]
    aud_in = [torch.empty_like(t) for t in aud]

    def h2d_copy():
        for d_, h_ in zip(vis_in + aud_in, vis_h + aud_h):
            d_.copy_(h_, non_blocking=True)

    def on_pair_streams(work):
        """Run work(entry_index) for the audio and the visual adapter of every position on two streams (fork/join)."""
        cur = torch.cuda.current_stream(dev)
        for i in range(0, len(stack), 2):
            s_aud.wait_stream(cur)
            s_vis.wait_stream(cur)
            for idx, strm in ((i, s_aud), (i + 1, s_vis)):
                with torch.cuda.stream(strm):
                    work(idx)
            cur.wait_stream(s_aud)
            cur.wait_stream(s_vis)

    def capture(fn, what):
        """CUDA-graph capture of fn() after warming it on a side stream (workspaces, packed weights, allocator)."""
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            ret = fn()
        return g, ret

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    # public-API forward (the call a user makes: 48 x VisualAdapter.forward on permuted views) over the static inputs
    last_maps = [None, None]
    last_out = [None]

    def api_forward():
        def one(idx):
            kind, s, p, pk = stack[idx]
            m = modules[idx]
            x, o = (vis_in[s], aud_in[s]) if kind == "vis" else (aud_in[s], vis_in[s])
            out, smap = m(x.permute(0, 2, 1).unsqueeze(-1), o.permute(0, 2, 1).unsqueeze(-1))
            last_maps[0 if kind == "aud" else 1] = smap
            last_out[0] = out
        with torch.no_grad():
            on_pair_streams(one)
            return torch.cat([last_maps[0].reshape(-1).float()[:BT * 64], last_maps[1].reshape(-1).float()[:BT * 36],
                              last_out[0].float().sum().reshape(1)])

    # The 48-call step is launch-bound when issued eagerly from Python (~1100 small launches), so it is captured once
    # into a CUDA graph and replayed (same kernels, same buffers; --no-graph times the eager loop instead).
    n0 = L.launch_count()
    for _ in range(2):
        step_resident()
    torch.cuda.synchronize(dev)
    launches_per_step = (L.launch_count() - n0) // 2
    run_step = step_resident
    launch_mode = "eager"
    if not args.no_graph:
        graph, _ = capture(step_resident, "fwd")
        run_step = graph.replay
        launch_mode = "cuda-graph replay of the 48-call step"
    fwd_ms = None
    if args.mode != "train":
        if not args.no_graph:
            g_api, res_static = capture(api_forward, "api fwd")

            def step_e2e(skip_h2d=False):
                if not skip_h2d:
                    h2d_copy()
                g_api.replay()
                d2h_buf[:res_static.numel()].copy_(res_static, non_blocking=True)
                return res_static.numel() * 4
        else:
            def step_e2e(skip_h2d=False):
                if not skip_h2d:
                    h2d_copy()
                res = api_forward()
                d2h_buf[:res.numel()].copy_(res, non_blocking=True)
                return res.numel() * 4
    else:
        from dgsct_b200.ddp import FlatAdam, GradBucketer
        fwd_ms = timed(run_step, max(3, args.steps // 2), 3)            # forward-only reference point (same stack)
        for m in modules:
            m.train()
            for n_, p_ in m.named_parameters():                        # declared, never used: gate_tk (net_trans.py:487), the dead
                if n_ == "gate_tk" or n_.startswith(("fc_caption", "temporal_gated")) or \
                        (CFG["resampler"] == "bicubic" and n_.startswith("conv_adapter")):     # caption / temporal heads, S4's conv_adapter
                    p_.requires_grad_(False)
        groups = [[p for p in m.parameters() if p.requires_grad] for m in reversed(modules)]   # backward completion order
        comm_dtype = torch.bfloat16 if (args.comm_dtype == "bf16" and world > 1) else torch.float32
        if args.torch_adam:     # round-1 configuration: torch's fused multi-tensor Adam over ~1400 parameter tensors
            bucketer = GradBucketer(groups, bucket_bytes=32 << 20, comm_dtype=comm_dtype)
            opt_adam = torch.optim.Adam([p for m in modules for p in m.parameters() if p.requires_grad], lr=1e-5, fused=True,
                                        capturable=not args.no_graph)
        else:                   # the engine's Adam: one kernel per flat bucket, reads the (bf16) reduced gradient in place
            bucketer = GradBucketer(groups, bucket_bytes=32 << 20, flatten_params=True, comm_dtype=comm_dtype)
            opt_adam = FlatAdam(bucketer, lr=1e-5, modules=modules)
        for t in vis_in + aud_in:
            t.requires_grad_(True)
        h2d_copy_nograd = h2d_copy

        def h2d_copy():      # noqa: F811  (leaf tensors: copy under no_grad)
            with torch.no_grad():
                h2d_copy_nograd()

        h2d_copy()
        ggen = torch.Generator(device=dev).manual_seed(7)
        d_vis = [torch.randn(t.shape, generator=ggen, device=dev, dtype=torch.float32).mul_(1e-2).to(dtype) for t in vis]
        d_aud = [torch.randn(t.shape, generator=ggen, device=dev, dtype=torch.float32).mul_(1e-2).to(dtype) for t in aud]
        dm_vis = [torch.randn(t.shape[0], 1, t.shape[1], generator=ggen, device=dev, dtype=torch.float32).to(dtype) for t in vis]
        dm_aud = [torch.randn(t.shape[0], 1, t.shape[1], generator=ggen, device=dev, dtype=torch.float32).to(dtype) for t in aud]
        nmod = len(modules)
        outs = [None] * nmod

        def fwd_one(idx):
            kind, s, p, pk = stack[idx]
            x, o = (vis_in[s], aud_in[s]) if kind == "vis" else (aud_in[s], vis_in[s])
            outs[idx] = modules[idx](x.permute(0, 2, 1).unsqueeze(-1), o.permute(0, 2, 1).unsqueeze(-1))

        def bwd_one(idx):
            kind, s, p, pk = stack[idx]
            out, smap = outs[idx]
            dgo = (d_vis[s] if kind == "vis" else d_aud[s]).permute(0, 2, 1).unsqueeze(-1)
            dgm = dm_vis[s] if kind == "vis" else dm_aud[s]
            torch.autograd.backward([out, smap], [dgo, dgm])

        train_streams = [2]

        def step_train():
            """forward through the 48 adapters (public nn.Module API), backward in reverse order with the gradient
            all-reduce launched per finished bucket, fused Adam.  With train_streams == 2 the audio and the visual
            adapter of one position (independent: they read the same two token tensors) run on two streams, forward
            and backward (autograd replays each node on its forward stream)."""
            cur = torch.cuda.current_stream(dev)
            bucketer.zero_grad()
            for t in vis_in + aud_in:
                t.grad = None
            if train_streams[0] == 2:
                on_pair_streams(fwd_one)
            else:
                for idx in range(nmod):
                    fwd_one(idx)
            for i in range(nmod - 2, -1, -2):                         # reverse position order
                if train_streams[0] == 2:
                    s_aud.wait_stream(cur)
                    s_vis.wait_stream(cur)
                    for idx, strm in ((i + 1, s_vis), (i, s_aud)):
                        with torch.cuda.stream(strm):
                            bwd_one(idx)
                    cur.wait_stream(s_aud)
                    cur.wait_stream(s_vis)
                else:
                    bwd_one(i + 1)
                    bwd_one(i)
                outs[i] = outs[i + 1] = None                          # drop the autograd graph of this position
                bucketer.group_done(nmod - 1 - (i + 1))
                bucketer.group_done(nmod - 1 - i)
            bucketer.finish()
            opt_adam.step()
            return torch.stack([vis_in[0].grad.float().abs().sum(), aud_in[0].grad.float().abs().sum()])

        for m in modules:
            m.fused_grad_accumulation = True      # kernels add straight into the bucket views (param.grad)
            m.save_activations = not args.recompute
        n0 = L.launch_count()
        step_train()
        torch.cuda.synchronize(dev)
        launches_per_step = L.launch_count() - n0
        run_step = step_train
        launch_mode = "eager"
        res_static = None
        if not args.no_graph:
            for ns in (2, 1):
                train_streams[0] = ns
                g_train, captured = None, 1
                try:
                    g_train, res_static = capture(step_train, "train")
                except Exception as e:                                # e.g. an op that cannot be captured
                    print(f"[bench] train-step graph capture with {ns} stream(s) failed ({type(e).__name__}: "
                          f"{str(e).splitlines()[0]})", file=sys.stderr)
                    torch.cuda.synchronize(dev)
                    captured = 0
                if dist is not None:                                  # every rank must take the same path: the warm-up
                    flag = torch.tensor([captured], device=dev)       # steps of an extra attempt run real collectives
                    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                    captured = int(flag.item())
                if captured:
                    run_step = g_train.replay
                    launch_mode = (f"cuda-graph replay of the whole train step (fwd + bwd + all-reduce + Adam), "
                                   f"{ns} stream(s)")
                    break
                run_step = step_train
                res_static = None
                g_train = None
                for k_ in range(nmod):
                    outs[k_] = None
        if res_static is None:
            train_streams[0] = 1                                      # eager: autograd stream syncs cost more than they hide

        def step_e2e(skip_h2d=False):      # noqa: F811  (host tokens in, fwd+bwd+all-reduce+Adam, two gradient checksums out)
            if not skip_h2d:
                h2d_copy()
            res = res_static
            if res_static is not None:
                run_step()
            else:
                res = step_train()
            d2h_buf[:2].copy_(res, non_blocking=True)
            return 8

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_step = timed(run_step, args.steps, args.warmup)
    launches = launches_per_step * args.steps
    clocks = sampler.stop() if sampler else None
    quick = os.environ.get("DGSCT_BENCH_QUICK") == "1"      # profiling runs under ncu: timed region only
    if quick:
        if os.environ.get("DGSCT_BENCH_PROFILE") == "1":    # ncu --profile-from-start off: exactly one eager step
            prof_step = step_train if args.mode == "train" else step_resident
            torch.cuda.synchronize(dev)
            torch.cuda.profiler.start()
            prof_step()
            torch.cuda.synchronize(dev)
            torch.cuda.profiler.stop()
        if rank == 0:
            print(json.dumps({"quick": True, "ms_per_step": ms_step, "clips_per_s": world * args.B / (ms_step * 1e-3)}), flush=True)
        return
    # ---- end to end through the public API: host tokens in (pinned -> device, every step), result out.  The copy of step
    # i+1's inputs is issued on a copy stream while step i computes (what a data loader with prefetch does): H2D lands in a
    # staging set, and the step starts with a device-to-device copy into the graph's static inputs.
    copy_stream = torch.cuda.Stream(dev)
    stage_in = [torch.empty_like(t) for t in vis_in + aud_in]
    static_in = vis_in + aud_in
    host_in = vis_h + aud_h
    prefetched = [False]

    def prefetch():
        copy_stream.wait_stream(torch.cuda.current_stream(dev))     # staging is free once the previous D2D copy was queued
        with torch.cuda.stream(copy_stream), torch.no_grad():
            for d_, h_ in zip(stage_in, host_in):
                d_.copy_(h_, non_blocking=True)
        prefetched[0] = True

    def step_e2e_prefetch():
        if not prefetched[0]:
            prefetch()
        torch.cuda.current_stream(dev).wait_stream(copy_stream)
        with torch.no_grad():
            for d_, s_ in zip(static_in, stage_in):
                d_.copy_(s_, non_blocking=True)
        prefetch()                                                   # next step's inputs travel under this step's compute
        return step_e2e(skip_h2d=True)

    d2h = step_e2e()
    ms_e2e_serial = timed(step_e2e, max(3, args.steps // 2), 3)
    ms_e2e = timed(step_e2e_prefetch, max(3, args.steps // 2), 3)

    # ---- per-stage device time (CUDA events inside the library, on the launching stream; one stream, side stream off)
    roof = None
    stage_ms = None
    eager_gpu = None

    def step_sequential():                # one stream, so per-stage event times do not overlap
        for kind, s_, p_, pk_ in stack:
            x_, o_ = (vis[s_], aud[s_]) if kind == "vis" else (aud[s_], vis[s_])
            adapter_forward_raw(CFG, p_, x_, o_, training=False, packed=pk_)

    prof_fn = step_sequential
    if args.mode == "train":
        train_streams[0] = 1
        prof_fn = step_train
    psteps = 2
    prof = None
    # The profiled train step contains the gradient all-reduce: EVERY rank must run it (rank 0 alone would wait for its
    # peers inside NCCL forever); only rank 0 keeps the stage times.  The forward-only profile has no collective.
    if rank == 0 or (args.mode == "train" and dist is not None):
        prof_fn()
        torch.cuda.synchronize(dev)
        L.profile_enable(True)
        for _ in range(psteps):
            prof_fn()
        torch.cuda.synchronize(dev)
        prof = L.profile_collect()
        L.profile_enable(False)
    if rank == 0:
        stage_ms = {k: round(v[0] / psteps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json: sustained bf16, copy bandwidth)" if peaks else "fallback"
        work = stage_work(BT, CFG)
        prof_total = max(sum(v[0] for v in prof.values()) / psteps, 1e-9)
        entries = {}
        for st_name, (amount, bound, what) in work.items():
            t_ms = prof.get(st_name, (0.0, 0))[0] / psteps
            if t_ms <= 0:
                continue
            peak = tf_peak if bound == "tensor" else hbm_peak
            ach = amount / (t_ms * 1e-3) / (1e12 if bound == "tensor" else 1e9)
            entries[st_name] = {"kernel": what, "bound": bound, "achieved": ach, "peak": peak,
                                "unit": "TFLOP/s" if bound == "tensor" else "GB/s", "frac": ach / peak, "ms_per_step": t_ms,
                                "launches_per_step": prof[st_name][1] // psteps, "share_of_step": t_ms / prof_total}
        # the dominant kernel of THIS timed step = the modelled stage with the largest device time
        dom = max(entries, key=lambda k: entries[k]["ms_per_step"]) if entries else None
        roof = dict(entries[dom]) if dom else {"kernel": None, "bound": "tensor", "achieved": 0.0, "peak": tf_peak, "unit": "TFLOP/s", "frac": 0.0}
        roof["stage"] = dom
        roof["peak_source"] = peak_src
        tr = load_traffic(dom)
        roof["traffic"] = tr["bytes"] if tr else None     # dram bytes per launch from a committed `ncu --set full` capture, else null
        if tr:
            roof["traffic_source"] = tr["source"]
            roof["traffic_algorithmic"] = tr.get("algorithmic_bytes")      # algorithmic bytes of that same launch
        roof["stages"] = entries
        tot_bytes = sum(alg_bytes(*dims(k, s), BT, K=CFG["K"], r=CFG["r"], g=CFG["g"]) for k, _, _, s in layer_list())
        tot_flops = sum(alg_flops(*dims(k, s), BT, K=CFG["K"], r=CFG["r"], g=CFG["g"]) for k, _, _, s in layer_list())
        t_fwd = (fwd_ms if fwd_ms is not None else ms_step) * 1e-3      # adapter-forward time (the metric's second half)
        roof["stack"] = {"what": "whole adapter-forward stack (48 calls), algorithmic bytes / flops of SURVEY.md 8(d)",
                         "algorithmic_GB": tot_bytes / 1e9, "algorithmic_TFLOP": tot_flops / 1e12, "fwd_ms": t_fwd * 1e3,
                         "hbm_GBps": tot_bytes / t_fwd / 1e9, "hbm_frac": tot_bytes / t_fwd / 1e9 / hbm_peak,
                         "tensor_TFLOPs": tot_flops / t_fwd / 1e12, "tensor_frac": tot_flops / t_fwd / 1e12 / tf_peak}
        if args.mode == "train":     # whole train step: forward + activation gradients + weight gradients ~ 3x forward flops
            roof["step"] = {"what": "whole train step, 3 x the forward's algorithmic flops", "algorithmic_TFLOP": 3 * tot_flops / 1e12,
                            "ms": ms_step, "tensor_TFLOPs": 3 * tot_flops / (ms_step * 1e-3) / 1e12,
                            "tensor_frac": 3 * tot_flops / (ms_step * 1e-3) / 1e12 / tf_peak}
        if world == 1 and not args.no_eager:
            eager_gpu = eager_gpu_comparator(args)

    if rank == 0:
        value = world * args.B / (ms_step * 1e-3)
        line = {"metric": metric_name(args), "value": value,
                "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "impl": "engine", "config": workload_config(args, torch.cuda.get_device_name(dev), launch_mode),
                "e2e": {"value": world * args.B / (ms_e2e * 1e-3), "unit": "clips/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e,
                        "how": "public nn.Module API; every step copies its pinned host tokens to the device (prefetched on a copy "
                               "stream under the previous step's compute, then one device-to-device copy) and reads the result back",
                        "serial_copy": {"value": world * args.B / (ms_e2e_serial * 1e-3), "ms_per_step": ms_e2e_serial,
                                        "how": "same, with the H2D copy serialised in front of the step"}},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
                "cpu_baseline": cpu_base if cpu_base is not None else {"value": None, "unit": "clips/s", "note": "skipped (--no-cpu-baseline)"},
                "gpu_eager_baseline": eager_gpu, "stage_ms_per_step": stage_ms}
        if fwd_ms is not None:
            line["forward_only"] = {"ms_per_step": fwd_ms, "clips_per_s": world * args.B / (fwd_ms * 1e-3)}
            line["collective"] = {"op": f"all-reduce (mean) of adapter gradients, {args.comm_dtype if world > 1 else 'fp32'} on the wire, "
                                        "fp32 master buckets of 32 MB, NCCL" if world > 1 else "none (1 GPU)",
                                  "bytes_per_step": bucketer.total_bytes() if world > 1 else 0}
            line["optimizer"] = "torch.optim.Adam(fused=True)" if args.torch_adam else "dgsct FlatAdam (one fused kernel per flat bucket)"
        print(json.dumps(line), flush=True)
    if dist is not None:
        # Captured graphs hold NCCL kernels: tearing the communicator down underneath them can block (seen at N=2: the
        # line above was printed and destroy_process_group never returned).  Drain, line up, and leave without it.
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def cpu_baseline(args):
    """Oracle on the host cores, bounded sample: ONE clip (BT=T) through the 48 adapters (same mode), best of 2-3 -- the same
    code the reference arm (--impl reference) times at the full B clips per step."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = make_cpu_step(args, 1)
    best = 1e30
    for it in range(4 if args.mode != "train" else 3):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if it:
            best = min(best, dt)
    return {"value": 1.0 / best, "unit": "clips/s", "cores": cores, "kind": "port", "clips_per_step": 1,
            "sample": f"1 clip (BT={args.T}) through the 48 adapters, fp32, {args.mode}, torch CPU {cores} threads, best of "
                      f"{3 if args.mode != 'train' else 2}; measured on rank 0 before the process group is created"}


def eager_gpu_comparator(args):
    """PyTorch eager on the same B200 ('the bar a hand-written kernel must beat'): the oracle's torch ops on
    cuda:0 under bf16 autocast, same workload, B clips per step.  A baseline leg like cpu_baseline (the oracle is never on
    the engine path); bounded to a few steps."""
    try:
        step = make_cpu_step(args, args.B, device="cuda", autocast=True)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        return {"value": args.B / (ms * 1e-3), "unit": "clips/s", "ms_per_step": ms, "kind": "pytorch-eager (oracle ops on cuda:0, bf16 autocast, "
                f"torch.optim.Adam), {args.mode}, B={args.B}", "steps": n}
    except Exception as e:       # informative only
        return {"value": None, "error": f"{type(e).__name__}: {str(e).splitlines()[0][:200]}"}


def stage_work(BT, cfg):
    """Algorithmic work per step of the library stages with a clean model: {stage: (flops or bytes, bound, description)}.
    Tensor-bound stages count EXECUTED flops (the re-associated resampler contracts over min(C, C') columns); HBM-bound row
    kernels count the bytes they must move (bf16 activations, fp32 up-sampler output)."""
    conv = dw = dwres = 0
    gate_ln = out_ln = bwd_out = bwd_gate_ln = tok_attn = x_attn = bwd_attn = 0
    K, r = cfg["K"], cfg["r"]
    sparse = cfg["resampler"] == "bicubic"
    for kind, _, _, s in layer_list():
        C, N, Co, No = dims(kind, s)
        Cd, H = C // r, C // 2
        cc = min(C, Co) if not sparse else Co
        conv += (2 * BT * N * No * cc) if not sparse else 0
        dwres += (2 * BT * N * No * cc) if not sparse else 0
        dw += 2 * BT * N * (2 * C * Cd + H * C + C * C + K * C) + 2 * BT * (No if (C < Co and not sparse) else N) * C * Co
        tok_attn += BT * N * C * 2 + BT * K * C * 6
        x_attn += BT * N * C * 4 + BT * N * 2 * ((K + 7) // 8 * 8) + BT * K * C * 2
        bwd_attn += BT * N * C * 10 + BT * N * 2 * ((K + 7) // 8 * 8) + BT * K * C * 16
        gate_ln += BT * N * C * 4
        out_ln += BT * N * C * 6
        bwd_out += BT * N * C * 8
        bwd_gate_ln += BT * N * C * 6
    w = {"tok_update": (tok_attn, "hbm", "fused token attention, tokens over V: softmax_n(T V^T), T' = T + P V in one pass over V "
                                        "(bulk-async staged tiles, warp-level MMAs; read V, write T')"),
         "x_update": (x_attn, "hbm", "fused token attention, stream over tokens: softmax_k(X T'^T), X' = X + g Q T' (read X, write X' and Q)"),
         "bwd_token": (bwd_attn, "hbm", "backward of the token attention, two fused kernels (read dX', X, V; write dX, dV; dT' in L2)"),
         "gate_ln": (gate_ln, "hbm", "gate combine + ln_before row kernel (read X', write Y)"),
         "out_ln": (out_ln, "hbm", "bn2 + ln_post + gate row kernel (read fp32 up-sampler output, write out)"),
         "bwd_out": (bwd_out, "hbm", "backward of gate / ln_post / bn2 sums (read d_out + Uo, write dUo)"),
         "bwd_gate_ln": (bwd_gate_ln, "hbm", "backward of ln_before + gate combine (read dY + X', write dX')"),
         "bwd_dw": (dw, "tensor", "weight-gradient products sum_b A_b^T B_b of the small matrices (tcgen05 GEMM, split over frames)")}
    if not sparse:
        w["resample_conv"] = (conv, "tensor", "token-resampler contraction (tcgen05 GEMM, 48 launches/step)")
        w["bwd_dother"] = (conv, "tensor", "d other = W_res^T . dU (tcgen05 GEMM)")
        w["bwd_dwres"] = (dwres, "tensor", "d W_res = sum_b dU_b . other_b^T (tcgen05 GEMM, K extended over frames)")
    return w


def load_traffic(stage):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the stage's kernel from the committed `ncu --set full`
    capture (profiles/r02_traffic.json: {stage: {"bytes": ..., "source": ...}}); None when no capture covers the stage."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        return t.get(stage)
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--B", type=int, default=0, help="clips per GPU (default: the workload's)")
    ap.add_argument("--T", type=int, default=0, help="frames per clip (default: the workload's)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--mode", default="train", choices=["fwd", "train"],
                    help="train: fwd + bwd + gradient all-reduce + Adam over the 48 adapters (BASELINE config 2/3); fwd: forward only")
    ap.add_argument("--no-graph", action="store_true", help="time the eager Python loop instead of CUDA-graph replay")
    ap.add_argument("--workload", default="ave", choices=sorted(WORKLOADS),
                    help="ave: BASELINE.json configs 1-3 (default, the metric's own); b224: Swin-V2-B/224 token shapes; "
                         "s4: config 4 (AVS-S4, B=4 x 5 frames); avqa_t60: config 5 (AVQA, T=60)")
    ap.add_argument("--comm-dtype", default="bf16", choices=["bf16", "f32"], help="dtype of the gradient all-reduce (N > 1)")
    ap.add_argument("--torch-adam", action="store_true", help="torch.optim.Adam(fused=True) instead of the engine's flat-bucket Adam")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager", action="store_true", help="skip the PyTorch-eager-on-GPU comparator (rank 0, N=1)")
    ap.add_argument("--ref-clips", type=int, default=0, help="reference arm: clips per step (default: the engine's B)")
    ap.add_argument("--recompute", action="store_true",
                    help="train: keep no activations, the backward recomputes the forward (default: saved activations)")
    args = ap.parse_args()
    _WL["name"] = args.workload
    args.B = args.B or wl()["B"]
    args.T = args.T or wl()["T"]
    if args.T != wl()["cfg"]["T"]:
        wl()["cfg"] = dict(wl()["cfg"], T=args.T)
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_engine(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
