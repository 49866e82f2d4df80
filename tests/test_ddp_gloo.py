"""world_size-2 gloo test of the gradient bucketer (the N>1 host logic; no GPU needed)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from dgsct_b200.ddp import GradBucketer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                   # identical replicas
    groups = [[torch.nn.Parameter(torch.randn(7, 5)), torch.nn.Parameter(torch.randn(5))],
              [torch.nn.Parameter(torch.randn(300, 40)), torch.nn.Parameter(torch.randn(3), requires_grad=False)],
              [torch.nn.Parameter(torch.randn(11))]]
    bk = GradBucketer(groups, bucket_bytes=1000)            # -> 3 buckets: [g0], [g1], [g2] (g0 < 1000 B merges with g1? no: 160+20 B then 48 KB)
    nb = len(bk.buckets)
    for step in range(2):
        bk.zero_grad()
        torch.manual_seed(100 * step + rank)               # rank-dependent data
        for gi in reversed(range(len(groups))):            # backward order
            for p in groups[gi]:
                if p.requires_grad:
                    g = torch.randn_like(p)
                    p.grad.add_(g)                          # autograd accumulates in place into the bucket view
            bk.group_done(gi)
        bk.finish()
        # expected: mean over ranks of the same random draws
        exp = []
        for r in range(world):
            torch.manual_seed(100 * step + r)
            cur = []
            for gi in reversed(range(len(groups))):
                for p in groups[gi]:
                    if p.requires_grad:
                        cur.append(torch.randn_like(p))
            exp.append(cur)
        k = 0
        ok = True
        for gi in reversed(range(len(groups))):
            for p in groups[gi]:
                if p.requires_grad:
                    mean = sum(e[k] for e in exp) / world
                    ok = ok and torch.allclose(p.grad, mean, atol=1e-6)
                    ok = ok and p.grad.data_ptr() >= bk.buckets[bk.owner[gi]].flat.data_ptr()
                    k += 1
        q.put((rank, step, ok, nb, bk.total_bytes()))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(4)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[2] for r in res), res
    assert res[0][3] >= 2 and res[0][4] == (36 + 8 + 12000 + 12) * 4      # every view starts on a 16-byte boundary


def _worker_ddp(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from dgsct_b200.ddp import AdapterDDP
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(10 + rank)                            # DIFFERENT initial replicas: construction must broadcast rank 0's
    mods = [torch.nn.Linear(6, 6), torch.nn.Linear(6, 6), torch.nn.Linear(6, 3)]
    mods[1].bias.requires_grad_(False)
    ddp = AdapterDDP(mods, bucket_bytes=64)
    w0 = [p.detach().clone() for m in mods for p in m.parameters()]
    gathered = [torch.zeros_like(w0[0]) for _ in range(world)]
    dist.all_gather(gathered, w0[0])
    same_init = all(torch.equal(gathered[0], g) for g in gathered)

    def loss_on(seed):
        torch.manual_seed(seed)
        x = torch.randn(4, 6)
        h = x
        for m in mods:
            h = torch.tanh(m(h))
        return (h ** 2).sum()

    # one synchronised step
    ddp.zero_grad()
    loss_on(100 + rank).backward()
    ddp.finish()
    got = [p.grad.clone() for m in mods for p in m.parameters() if p.requires_grad]
    # reference: plain autograd per rank, averaged with explicit all-reduces
    ref_mods = [torch.nn.Linear(6, 6), torch.nn.Linear(6, 6), torch.nn.Linear(6, 3)]
    for a, b in zip(ref_mods, mods):
        a.load_state_dict(b.state_dict())
    ref_mods[1].bias.requires_grad_(False)
    torch.manual_seed(100 + rank)
    h = torch.randn(4, 6)
    for m in ref_mods:
        h = torch.tanh(m(h))
    (h ** 2).sum().backward()
    ok = same_init
    k = 0
    for m in ref_mods:
        for p in m.parameters():
            if p.requires_grad:
                g = p.grad.clone()
                dist.all_reduce(g)
                ok = ok and torch.allclose(got[k], g / world, atol=1e-6)
                k += 1
    # gradient accumulation: first micro-batch without collectives, second with
    ddp.zero_grad()
    with ddp.no_sync():
        loss_on(200 + rank).backward()
    loss_on(300 + rank).backward()
    ddp.finish()
    acc = [p.grad.clone() for m in mods for p in m.parameters() if p.requires_grad]
    for m in ref_mods:
        m.zero_grad()
    for seed in (200 + rank, 300 + rank):
        torch.manual_seed(seed)
        h = torch.randn(4, 6)
        for m in ref_mods:
            h = torch.tanh(m(h))
        (h ** 2).sum().backward()
    k = 0
    for m in ref_mods:
        for p in m.parameters():
            if p.requires_grad:
                g = p.grad.clone()
                dist.all_reduce(g)
                ok = ok and torch.allclose(acc[k], g / world, atol=1e-6)
                k += 1
    q.put((rank, bool(ok), len(ddp.bucketer.buckets)))
    dist.destroy_process_group()


def test_adapter_ddp_hooks_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_ddp, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert res[0][2] >= 2


# ---- engine-style modules: gradients are added into the bucket views from INSIDE the autograd function ---------------
class _FakeEngineFn(torch.autograd.Function):
    """Stands in for dgsct_b200.adapter._AdapterFn on the CPU: the 'kernels' add the weight gradient straight into
    ``param.grad`` (a bucket view), autograd gets None for it, and the DDP completion callback runs at the end."""

    @staticmethod
    def forward(ctx, module, x, w):
        ctx.module = module
        ctx.save_for_backward(x)
        return x @ w.detach()

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        m = ctx.module
        m.order.append("fn_backward")
        m.weight.grad.add_(x.t() @ g)
        cb = getattr(m, "_ddp_done_cb", None)
        if cb is not None:
            cb()
        return None, g @ m.weight.detach().t(), None


class _FakeEngine(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.randn(6, 6))
        self.order = []

    def _packed_weights(self, *a):      # marks the module as an engine adapter for AdapterDDP
        return None

    def invalidate_packed(self):
        self.order.append("invalidate")

    def forward(self, x):
        return _FakeEngineFn.apply(self, x, self.weight)


def _worker_engine(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from dgsct_b200.ddp import AdapterDDP
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(5 + rank)
    first, second, unused = _FakeEngine(), _FakeEngine(), torch.nn.Linear(6, 2)
    ddp = AdapterDDP([first, second, unused], bucket_bytes=16)          # one bucket per module
    ok = "invalidate" in first.order                                     # broadcast went through .data: cache invalidated
    W1, W2 = first.weight.detach().clone(), second.weight.detach().clone()

    def run():
        torch.manual_seed(50 + rank)
        x = torch.randn(4, 6)                   # does NOT require grad (frozen patch_embed output, net_trans.py:877,891)
        h = second(first(x)) + second(first(2 * x))     # every module is invoked TWICE per step
        (h ** 2).sum().backward()

    def expect():
        outs = []
        for r in range(world):
            torch.manual_seed(50 + r)
            x = torch.randn(4, 6)
            a, b = W1.clone().requires_grad_(True), W2.clone().requires_grad_(True)
            (((x @ a) @ b + ((2 * x) @ a) @ b) ** 2).sum().backward()
            outs.append((a.grad, b.grad))
        return sum(o[0] for o in outs) / world, sum(o[1] for o in outs) / world

    e1, e2 = expect()
    ddp.zero_grad()
    run()
    ddp.finish()                                # the unused Linear never reports: finish() must still reduce its bucket
    ok = ok and torch.allclose(first.weight.grad, e1, atol=1e-5) and torch.allclose(second.weight.grad, e2, atol=1e-5)
    ok = ok and float(unused.weight.grad.abs().max()) == 0.0
    # the reference loop's optimizer.zero_grad() (set_to_none=True) drops the bucket views: they are re-attached
    ddp.zero_grad()
    for m in (first, second, unused):
        for p in m.parameters():
            p.grad = None
    ddp.zero_grad()
    ok = ok and first.weight.grad is not None and first.weight.grad.data_ptr() == ddp.bucketer.buckets[ddp.bucketer.owner[2]].flat.data_ptr()
    run()
    ddp.finish()
    ok = ok and torch.allclose(first.weight.grad, e1, atol=1e-5) and torch.allclose(second.weight.grad, e2, atol=1e-5)
    # ordinary module whose gradient was produced OUTSIDE the bucket (grad reset to None after zero_grad): folded in by finish
    ddp.zero_grad()
    unused.weight.grad = None
    unused.bias.grad = None
    torch.manual_seed(70 + rank)
    xin = torch.randn(3, 6)
    (unused(xin) ** 2).sum().backward()
    ddp.finish()
    g = torch.autograd.grad((torch.nn.functional.linear(xin, unused.weight, unused.bias) ** 2).sum(), unused.weight)[0]
    dist.all_reduce(g)
    ok = ok and torch.allclose(unused.weight.grad, g / world, atol=1e-5)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_engine_modules_report_from_inside_backward_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_engine, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res


def _worker_bf16(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from dgsct_b200.ddp import GradBucketer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    groups = [[torch.nn.Parameter(torch.randn(33, 7))], [torch.nn.Parameter(torch.randn(5))]]
    w0 = [p.detach().clone() for g in groups for p in g]
    bk = GradBucketer(groups, bucket_bytes=64, flatten_params=True, comm_dtype=torch.bfloat16)
    ok = all(torch.equal(p.detach(), w) for p, w in zip([p for g in groups for p in g], w0))        # values survive flattening
    ok = ok and all(p.data_ptr() % 16 == 0 and p.grad.data_ptr() % 16 == 0 for g in groups for p in g)
    bk.zero_grad()
    for gi in (1, 0):
        for p in groups[gi]:
            p.grad.add_(torch.full_like(p, float(rank + 1)))
        bk.group_done(gi)
    bk.finish()
    ok = ok and all(torch.allclose(p.grad, torch.full_like(p, 1.5), atol=1e-2) for g in groups for p in g)   # mean of 1 and 2, bf16 wire
    q.put((rank, bool(ok), bk.total_bytes()))
    dist.destroy_process_group()


def test_bf16_wire_and_flat_parameters_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_bf16, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert res[0][2] == (232 + 8) * 2
