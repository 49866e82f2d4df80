"""GPU comparator: the same adapter math as stock PyTorch eager kernels on the B200 -- the oracle's
torch ops moved to cuda:0 (cuBLAS bmm / linear, cuDNN-free elementwise) -- against the engine, on the bench workload
(48 calls, B=8, T=10).  The oracle is used here as the comparator only; the engine path never touches it."""
import importlib.util
import json
import os

import pytest
import torch

from dgsct_b200.adapter import adapter_forward_raw
from oracle import dgsct_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _time(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def test_pytorch_eager_comparator_on_the_bench_workload():
    b = _bench()
    dev = torch.device("cuda", 0)
    gen = torch.Generator().manual_seed(43)
    vis, aud = b.make_tokens(8, 10, gen, torch.float32)
    vis = [t.to(dev) for t in vis]
    aud = [t.to(dev) for t in aud]
    stack = []
    for kind, li, pos, s in b.layer_list():
        C, N, Co, No = b.dims(kind, s)
        p = {k: v.to(dev) for k, v in b.make_params(C, N, Co, No, gen).items()}
        stack.append((kind, s, O.make_config("ave", C, N, Co, No), p))

    def eager(dtype):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == torch.bfloat16):
            for kind, s, cfg, p in stack:
                x, o = (vis[s], aud[s]) if kind == "vis" else (aud[s], vis[s])
                O.adapter_forward(cfg, p, x, o, training=False)

    vis16, aud16 = [t.bfloat16() for t in vis], [t.bfloat16() for t in aud]
    packed = []
    for kind, s, cfg, p in stack:
        x, o = (vis16[s], aud16[s]) if kind == "vis" else (aud16[s], vis16[s])
        packed.append(adapter_forward_raw(b.ADAPTER_CFG, p, x, o, pack_only=True))

    def engine():
        for (kind, s, cfg, p), pk in zip(stack, packed):
            x, o = (vis16[s], aud16[s]) if kind == "vis" else (aud16[s], vis16[s])
            adapter_forward_raw(b.ADAPTER_CFG, p, x, o, training=False, packed=pk)

    try:
        eager_ms = {"pytorch_eager_fp32_ms": _time(lambda: eager(torch.float32)),
                    "pytorch_eager_bf16_autocast_ms": _time(lambda: eager(torch.bfloat16))}
    except Exception as e:                      # the comparator is informative; it must never fail the parity suite
        pytest.skip(f"oracle comparator did not run on CUDA: {type(e).__name__}: {e}")
    ms = dict(eager_ms, engine_bf16_eager_launch_ms=_time(engine))
    ms["clips_per_s"] = {k[:-3]: 8.0 / (v * 1e-3) for k, v in ms.items()}
    print("COMPARATOR", json.dumps(ms))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        json.dump(ms, open(os.path.join(out, "eager_comparator.json"), "w"))
    # recorded, not enforced: bench.py's graph-replayed step (not this eager launch loop) is the engine's headline
    assert ms["engine_bf16_eager_launch_ms"] > 0 and ms["pytorch_eager_bf16_autocast_ms"] > 0
