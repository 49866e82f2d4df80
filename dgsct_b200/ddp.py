"""Data-parallel training support for the adapter engine: one process per GPU, clips sharded over
ranks, adapters replicated, ONE collective per optimizer step -- the all-reduce (mean) of the trainable-parameter
gradients over NCCL/NVLink, bucketed and launched as each bucket's backward finishes so it overlaps the backward of
the remaining adapters.  (The reference has no DDP: AVE/AVVP are single-GPU, the others use nn.DataParallel; BatchNorm
statistics stay per rank exactly as DataParallel replicas do.)

The bucketer is backend-agnostic (NCCL on GPUs, gloo in the CPU tests): it owns flat fp32 buffers and makes every
``param.grad`` a view into its bucket, so gradients are produced in place and no flatten/unflatten copies exist.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


class GradBucket:
    """One flat fp32 gradient buffer; every ``param.grad`` is a view into it (16-byte aligned offsets).
    ``flatten_params``: the parameters themselves become views of a second flat buffer with the same layout, so an
    optimizer can update the whole bucket with one elementwise kernel (:class:`FlatAdam`).
    ``comm_dtype=torch.bfloat16``: the all-reduce runs on a bf16 copy of the bucket (half the NVLink bytes)."""

    def __init__(self, params: Sequence[torch.nn.Parameter], flatten_params: bool = False, comm_dtype: torch.dtype = torch.float32):
        self.params = list(params)
        self.offsets: List[int] = []
        n = 0
        for p in self.params:
            self.offsets.append(n)
            n += _pad4(p.numel())
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views: List[torch.Tensor] = []
        for p, off in zip(self.params, self.offsets):
            if p.dtype != torch.float32:
                raise TypeError("gradient buckets hold fp32 master gradients")
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            p.grad = self.views[-1]
        self.pflat: Optional[torch.Tensor] = None
        if flatten_params:
            self.pflat = torch.zeros(n, dtype=torch.float32, device=dev)
            with torch.no_grad():
                for p, off in zip(self.params, self.offsets):
                    dst = self.pflat[off:off + p.numel()].view_as(p)
                    dst.copy_(p.data)
                    p.data = dst                      # same Parameter object, storage now inside the flat buffer
        self.comm_dtype = comm_dtype
        self.comm: Optional[torch.Tensor] = None      # bf16 communication copy (filled when the bucket is reduced)
        if comm_dtype == torch.bfloat16:
            self.comm = torch.zeros(n, dtype=torch.bfloat16, device=dev)
        elif comm_dtype != torch.float32:
            raise TypeError("comm_dtype must be float32 or bfloat16")
        self.reduced_in_comm = False                  # this step's reduced SUM lives in `comm` (bf16), not in `flat`
        self.work: Optional[dist.Work] = None
        self.launched = False

    def reattach(self) -> None:
        """Make every ``param.grad`` alias its bucket view again.  ``optimizer.zero_grad()`` defaults to
        ``set_to_none=True`` (the reference loop calls it, main_trans.py:134), which drops the views: autograd would then
        allocate fresh gradients and the bucket would reduce stale zeros.  A foreign gradient tensor found here (produced
        after such a reset) is folded into the bucket before the view is restored."""
        for p, v in zip(self.params, self.views):
            g = p.grad
            if g is None:
                p.grad = v
            elif g.data_ptr() != v.data_ptr():
                v.add_(g.detach().to(v.dtype).view_as(v))
                p.grad = v

    def zero(self) -> None:
        self.flat.zero_()
        self.launched = False
        self.reduced_in_comm = False
        for p, v in zip(self.params, self.views):
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v

    def all_reduce(self, group=None, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Launch the averaging all-reduce of this bucket (asynchronously).  On CUDA the collective runs on ``stream``
        after the work already queued on the current stream, so later kernels of the backward overlap it."""
        self.launched = True
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return

        def reduce_():
            buf = self.flat
            if self.comm is not None:
                self._to_comm()
                buf = self.comm
                self.reduced_in_comm = True
            self.work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group, async_op=True)

        if stream is not None and self.flat.is_cuda:
            stream.wait_stream(torch.cuda.current_stream(self.flat.device))
            with torch.cuda.stream(stream):
                reduce_()
        else:
            reduce_()

    def _to_comm(self) -> None:
        if self.flat.is_cuda:
            import ctypes as C
            from . import _lib as L
            st = torch.cuda.current_stream(self.flat.device).cuda_stream
            with torch.cuda.device(self.flat.device):
                L.check(L.lib().dgsct_cast_bucket_bf16(C.c_void_p(self.flat.data_ptr()), C.c_void_p(self.comm.data_ptr()),
                                                       C.c_size_t(self.flat.numel()), C.c_float(1.0), C.c_void_p(st)), "cast_bucket")
        else:
            self.comm.copy_(self.flat)

    def wait(self, group=None, keep_in_comm: bool = False) -> None:
        """Wait for the all-reduce and turn the sum into the mean in ``flat``.  ``keep_in_comm`` (FlatAdam): leave a bf16
        reduced SUM in ``comm`` -- the optimizer kernel reads it directly with grad_scale = 1/world, saving one pass."""
        if self.work is not None:
            self.work.wait()
            self.work = None
            w = dist.get_world_size(group)
            if self.reduced_in_comm:
                if not keep_in_comm:
                    self.flat.copy_(self.comm)
                    self.flat.div_(w)
                    self.reduced_in_comm = False
            else:
                self.flat.div_(w)


class GradBucketer:
    """Groups parameters into buckets of about ``bucket_bytes`` in registration order (call it with the adapters in the
    order their backward completes, i.e. reversed forward order) and wires ``param.grad`` into the flat buffers.
    Parameters that never receive a gradient (``gate_tk``, dead ``fc_caption`` / ``temporal_gated`` heads) must be left
    out -- the all-reduce would still be correct (zeros), just wasteful."""

    def __init__(self, param_groups: Iterable[Sequence[torch.nn.Parameter]], bucket_bytes: int = 32 << 20, group=None,
                 flatten_params: bool = False, comm_dtype: torch.dtype = torch.float32):
        self.group = group
        self.keep_in_comm = False             # set by FlatAdam: bf16-reduced sums stay in the communication buffers
        mk = lambda ps: GradBucket(ps, flatten_params=flatten_params, comm_dtype=comm_dtype)
        self.buckets: List[GradBucket] = []
        self.owner: List[int] = []            # bucket index of every registered group, in order
        cur: List[torch.nn.Parameter] = []
        cur_bytes = 0
        pending_groups = 0
        for g in param_groups:
            g = [p for p in g if p.requires_grad]
            cur += g
            cur_bytes += sum(p.numel() for p in g) * 4
            pending_groups += 1
            if cur_bytes >= bucket_bytes:
                self.buckets.append(mk(cur))
                self.owner += [len(self.buckets) - 1] * pending_groups
                cur, cur_bytes, pending_groups = [], 0, 0
        if cur:
            self.buckets.append(mk(cur))
            self.owner += [len(self.buckets) - 1] * pending_groups
        self._remaining = [self.owner.count(i) for i in range(len(self.buckets))]
        self._left = list(self._remaining)
        dev = self.buckets[0].flat.device if self.buckets else torch.device("cpu")
        self.comm_stream = torch.cuda.Stream(dev) if dev.type == "cuda" else None
        self._comm_used = False

    def zero_grad(self) -> None:
        for b in self.buckets:
            b.zero()
        self._left = list(self._remaining)

    def group_done(self, group_index: int) -> None:
        """Tell the bucketer that the backward of registered group ``group_index`` has been enqueued; once every group of
        a bucket is done its all-reduce is launched."""
        bi = self.owner[group_index]
        self._left[bi] -= 1
        if self._left[bi] == 0:
            self._launch(bi)

    def _launch(self, bi: int) -> None:
        b = self.buckets[bi]
        if b.launched:
            raise RuntimeError(f"gradient bucket {bi} was already reduced in this step (a module ran its backward more often "
                               "than announced; call zero_grad() once per step)")
        b.reattach()
        b.all_reduce(self.group, self.comm_stream)
        self._comm_used = self._comm_used or b.work is not None

    def finish(self) -> None:
        """Reduce every bucket that has not been reduced yet (an adapter that was not used in this step, or a parameter
        that received no gradient, never reports: its bucket would otherwise keep per-rank gradients), wait for every
        outstanding all-reduce and turn sums into means (call before optimizer.step())."""
        for bi, b in enumerate(self.buckets):
            if not b.launched:
                self._launch(bi)
        for b in self.buckets:
            b.wait(self.group, keep_in_comm=self.keep_in_comm)
        # join the communication stream only if a collective was launched on it this step: with one rank nothing was,
        # and waiting on a stream that is not part of an ongoing CUDA-graph capture would invalidate the capture
        if self.comm_stream is not None and self._comm_used:
            torch.cuda.current_stream(self.comm_stream.device).wait_stream(self.comm_stream)
        self._comm_used = False

    def total_bytes(self) -> int:
        return sum(b.flat.numel() * (2 if b.comm is not None else 4) for b in self.buckets)


class FlatAdam:
    """torch.optim.Adam semantics (the reference's optimizer, main_trans.py:276) over the flat buckets of a
    :class:`GradBucketer` built with ``flatten_params=True``: ONE fused kernel per bucket (dgsct_adam_flat) instead of a
    multi-tensor apply over ~1400 parameter tensors, reading the (possibly bf16, all-reduced-SUM) gradient in place.
    The step counter lives on the device, so ``step()`` is CUDA-graph capturable.  After a step the engine adapters'
    packed bf16 weight copies are stale: pass the modules so they are invalidated (their next forward re-packs)."""

    def __init__(self, bucketer: GradBucketer, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 modules: Sequence[torch.nn.Module] = ()):
        self.bucketer = bucketer
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.modules = list(modules)
        for b in bucketer.buckets:
            if b.pflat is None:
                raise ValueError("FlatAdam needs GradBucketer(..., flatten_params=True)")
            if not b.flat.is_cuda:
                raise RuntimeError("FlatAdam runs on CUDA buckets only (no CPU fallback)")
        dev = bucketer.buckets[0].flat.device
        self.m = [torch.zeros_like(b.flat) for b in bucketer.buckets]
        self.v = [torch.zeros_like(b.flat) for b in bucketer.buckets]
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        bucketer.keep_in_comm = True

    def step(self) -> None:
        import ctypes as C
        from . import _lib as L
        lib = L.lib()
        world = dist.get_world_size(self.bucketer.group) if (dist.is_available() and dist.is_initialized()) else 1
        dev = self.step_dev.device
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            L.check(lib.dgsct_adam_tick(C.c_void_p(self.step_dev.data_ptr()), st), "adam_tick")
            for b, m, v in zip(self.bucketer.buckets, self.m, self.v):
                if b.reduced_in_comm:       # bf16 SUM over ranks, still in the communication buffer
                    g, g_bf16, scale = b.comm, 1, 1.0 / world
                    b.reduced_in_comm = False
                else:                       # fp32 mean (or a single rank's gradient)
                    g, g_bf16, scale = b.flat, 0, 1.0
                L.check(lib.dgsct_adam_flat(C.c_void_p(b.pflat.data_ptr()), C.c_void_p(g.data_ptr()), g_bf16, C.c_void_p(m.data_ptr()),
                                            C.c_void_p(v.data_ptr()), C.c_size_t(b.flat.numel()), C.c_float(self.lr),
                                            C.c_float(self.betas[0]), C.c_float(self.betas[1]), C.c_float(self.eps),
                                            C.c_float(self.weight_decay), C.c_float(scale), C.c_void_p(self.step_dev.data_ptr()), st),
                        "adam_flat")
        for mod in self.modules:
            if hasattr(mod, "invalidate_packed"):
                mod.invalidate_packed()


class AdapterDDP:
    """Data-parallel wrapper for the trainable adapter modules of a model (what the reference's
    single-GPU ``main()`` lacks, main_trans.py:178-302).  The frozen backbones need nothing; every listed module gets

    * its parameters (and buffers) broadcast from rank 0 at construction, so replicas start identical;
    * its gradients wired into the flat buckets of a :class:`GradBucketer` (registration = reverse forward order, the
      order in which backward finishes);
    * hooks that launch a bucket's all-reduce as soon as the backward of its last module has been enqueued:
      engine adapters (``fused_grad_accumulation``: the kernels add into ``param.grad`` and autograd never sees the
      gradient) use a module backward hook, ordinary modules use per-parameter post-accumulate hooks.

    Usage::

        ddp = AdapterDDP([m for m in model.modules() if isinstance(m, dgsct_b200.VisualAdapter)])
        for batch in loader:
            ddp.zero_grad(); loss = step(batch); loss.backward(); ddp.finish(); optimizer.step()

    ``with ddp.no_sync():`` runs a micro-batch of gradient accumulation (``--accum_itr``, main_trans.py:135-136) without
    collectives; gradients keep adding up in the buckets."""

    def __init__(self, modules: Sequence[torch.nn.Module], bucket_bytes: int = 32 << 20, group=None, fused: bool = True,
                 broadcast: bool = True, flatten_params: bool = False, comm_dtype: torch.dtype = torch.float32):
        self.modules = list(modules)
        self.group = group
        order = list(reversed(self.modules))
        if broadcast and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            with torch.no_grad():
                for m in self.modules:
                    for t in list(m.parameters()) + list(m.buffers()):
                        dist.broadcast(t.data, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
                    if hasattr(m, "invalidate_packed"):
                        m.invalidate_packed()        # .data writes do not bump _version (packed bf16 weight cache key)
        groups = [[p for p in m.parameters() if p.requires_grad] for m in order]
        self.bucketer = GradBucketer(groups, bucket_bytes=bucket_bytes, group=group, flatten_params=flatten_params,
                                     comm_dtype=comm_dtype)
        self._sync = True
        self._npar = [len(g) for g in groups]
        self._pending = list(self._npar)
        self._handles = []
        self._engine = []
        self._fwd_calls = [0] * len(order)       # engine adapters: forward calls of this step that will run a backward
        self._bwd_calls = [0] * len(order)
        for j, (m, g) in enumerate(zip(order, groups)):
            engine = fused and hasattr(m, "_packed_weights")          # dgsct_b200.VisualAdapter
            self._engine.append(engine)
            if engine:
                # The engine's kernels add into the bucket views from inside _AdapterFn.backward, so "this module's
                # gradients are complete" is signalled from the END of that function (module._ddp_done_cb).  A module
                # full-backward hook is NOT usable: when no module input requires grad (layer-0 adapters fed by the frozen
                # patch_embed, net_trans.py:877,891) PyTorch fires it before the backward function has run.
                m.fused_grad_accumulation = True
                m._ddp_done_cb = self._engine_done(j)
                self._handles.append(m.register_forward_pre_hook(self._count_forward(j)))
            else:
                for p in g:
                    self._handles.append(p.register_post_accumulate_grad_hook(self._param_hook(j)))

    def _count_forward(self, j):
        def hook(module, args):
            if torch.is_grad_enabled():
                self._fwd_calls[j] += 1
        return hook

    def _engine_done(self, j):
        def done():
            self._bwd_calls[j] += 1
            # a module called several times per step reports once, after the backward of its last call
            if self._sync and self._bwd_calls[j] >= max(self._fwd_calls[j], 1):
                self.bucketer.group_done(j)
        return done

    def _param_hook(self, j):
        def hook(param):
            if not self._sync:
                return
            self._pending[j] -= 1
            if self._pending[j] == 0:
                self.bucketer.group_done(j)
        return hook

    def zero_grad(self) -> None:
        self.bucketer.zero_grad()
        self._pending = list(self._npar)
        self._fwd_calls = [0] * len(self._fwd_calls)
        self._bwd_calls = [0] * len(self._bwd_calls)

    def no_sync(self):
        import contextlib

        @contextlib.contextmanager
        def ctx():
            prev, self._sync = self._sync, False
            try:
                yield
            finally:
                self._sync = prev
        return ctx()

    def finish(self) -> None:
        """Wait for the outstanding all-reduces; ``param.grad`` then holds the mean over ranks."""
        self.bucketer.finish()
        self._pending = list(self._npar)
        self._fwd_calls = [0] * len(self._fwd_calls)
        self._bwd_calls = [0] * len(self._bwd_calls)

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []
        for m in self.modules:
            if hasattr(m, "_ddp_done_cb"):
                m._ddp_done_cb = None
