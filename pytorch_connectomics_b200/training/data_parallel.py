"""The DDP seam as an OBJECT: a ``DistributedDataParallel``-shaped wrapper over the flat gradient arena.

The reference wraps the LightningModule in ``torch.nn.parallel.DistributedDataParallel`` through Lightning's ``DDPStrategy``
(``connectomics/training/lightning/trainer.py:231-256``: ``find_unused_parameters=True`` for MedNeXt) and hands
``accumulate_grad_batches`` / ``gradient_clip_val`` to ``pl.Trainer`` (``:314-334``).  What that gives the training loop:

* every rank starts from rank 0's parameters and buffers;
* ``loss.backward()`` alone leaves the MEAN gradient over ranks in ``p.grad`` — bucket all-reduces are launched while the rest
  of the backward pass still runs;
* parameters that receive no gradient (unused deep-supervision heads, ``dummy_tensor``) do not hang the exchange;
* inside ``no_sync()`` (non-boundary micro-batches of an accumulation window) nothing is exchanged;
* ``register_comm_hook(state, hook)`` replaces the exchange of one bucket.

:class:`ArenaDataParallel` gives the same contract on the B200 design: gradients live in ONE flat fp32 arena
(:class:`~.ddp.FlatGradArena`), which is cut into contiguous SEGMENTS in reverse parameter order (the order gradients become
ready).  A ``post_accumulate_grad`` hook marks a parameter ready; when the last parameter of a segment is ready — and every
earlier segment has been launched, so all ranks issue collectives in the same order — the segment's slice of the arena is
all-reduced asynchronously (NCCL runs it on its own stream over NVLink while the remaining weight-gradient kernels run).  A
callback queued on the autograd engine finishes the step at the end of ``backward()``: segments that never completed (their
parameters were unused this step: the arena slice is zero, the ``find_unused_parameters`` semantics without a graph walk) are
launched, every future is waited for on the current stream, and the result is the DDP mean — or the plain SUM with
``reduce_op="sum"``, in which case the 1/world is folded into the fused optimizer kernel (``FusedAdamW.step(grads_are_summed=
True)``) and the gradients are never touched by a separate scaling pass.

The hook protocol is DDP's: ``hook(state, bucket) -> torch.futures.Future[Tensor]`` where ``bucket`` offers ``buffer()``,
``index()``, ``is_last()``, ``parameters()``, ``gradients()``, ``set_buffer()`` (:class:`GradSegment` here,
``torch.distributed.GradBucket`` under the real DDP), so :func:`allreduce_sum_hook` / :func:`allreduce_mean_hook` /
:func:`bf16_compress_hook` work with ``DistributedDataParallel.register_comm_hook`` as well.
"""

from __future__ import annotations

import contextlib
from typing import Any, Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

from .ddp import FlatGradArena, broadcast_parameters

__all__ = ["ArenaDataParallel", "GradSegment", "allreduce_mean_hook", "allreduce_sum_hook", "bf16_compress_hook",
           "plan_segments", "make_arena_ddp_strategy"]


def _world(group) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def plan_segments(numels: Sequence[int], offsets: Sequence[int], total: int, cap_elems: int,
                  first_cap_elems: Optional[int] = None) -> List[tuple]:
    """Cut parameters ``0..n-1`` (arena order) into contiguous groups walking from the LAST parameter to the first — autograd
    produces gradients roughly in reverse registration order, which is also how DDP fills its buckets.  Returns
    ``[(first_param, last_param_exclusive, lo, hi)]`` in LAUNCH order; ``[lo, hi)`` is the arena element range, padding behind
    a parameter belongs to its segment.  A group closes once it holds at least ``cap_elems`` elements (the first one
    ``first_cap_elems``: DDP's small first bucket gets the exchange started early)."""
    n = len(numels)
    if n == 0:
        return []
    cap_elems = max(1, int(cap_elems))
    cap = max(1, int(first_cap_elems)) if first_cap_elems is not None else cap_elems
    out, hi_p, acc = [], n, 0
    for i in range(n - 1, -1, -1):
        acc += int(numels[i])
        if acc >= cap or i == 0:
            hi = int(offsets[hi_p]) if hi_p < n else int(total)
            out.append((i, hi_p, int(offsets[i]), hi))
            hi_p, acc, cap = i, 0, cap_elems
    return out


class GradSegment:
    """One contiguous slice of the gradient arena — the ``GradBucket`` a comm hook receives."""

    def __init__(self, arena: FlatGradArena, index: int, first: int, last: int, lo: int, hi: int, is_last: bool) -> None:
        self._arena, self._index, self.first, self.last, self.lo, self.hi, self._is_last = arena, index, first, last, lo, hi, is_last

    def index(self) -> int:
        return self._index

    def is_last(self) -> bool:
        return self._is_last

    def buffer(self) -> torch.Tensor:
        return self._arena.buffer[self.lo:self.hi]

    def set_buffer(self, tensor: torch.Tensor) -> None:
        buf = self.buffer()
        if tensor.data_ptr() != buf.data_ptr():
            buf.copy_(tensor.reshape(-1))

    def parameters(self) -> List[torch.Tensor]:
        return list(self._arena.params[self.first:self.last])

    def gradients(self) -> List[torch.Tensor]:
        return [self._arena.view_of(i) for i in range(self.first, self.last)]

    def __repr__(self) -> str:
        return f"GradSegment({self._index}: params [{self.first},{self.last}) arena [{self.lo},{self.hi}))"


# --------------------------------------------------------------------------------------------------------------------
# comm hooks (DDP's protocol; ``state`` is the process group or None).  No annotations on purpose: torch's DDP compares a
# hook's annotations with the real ``Future[Tensor]`` / ``GradBucket`` types, and this module's are strings.

def _group_of(state: Any):
    if state is None or isinstance(state, dist.ProcessGroup):
        return state
    return getattr(state, "process_group", None)


def allreduce_sum_hook(state, bucket):
    """SUM only: pair with ``FusedAdamW.step(grads_are_summed=True)`` (the 1/world rides in the optimizer kernel)."""
    buf = bucket.buffer()
    return dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=_group_of(state), async_op=True).get_future().then(
        lambda f: f.value()[0])


def allreduce_mean_hook(state, bucket):
    """What DDP does by default: SUM, then 1/world."""
    group = _group_of(state)
    inv = 1.0 / _world(group)
    buf = bucket.buffer()
    return dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group, async_op=True).get_future().then(
        lambda f: f.value()[0].mul_(inv))


def bf16_compress_hook(state, bucket):
    """Halve the bytes on the wire (torch's ``bf16_compress_hook`` semantics: divide first, reduce in bf16, decompress into the
    bucket).  Not bit-comparable with the fp32 exchange; for bandwidth-bound multi-node runs."""
    group = _group_of(state)
    buf = bucket.buffer()
    small = (buf / _world(group)).to(torch.bfloat16)

    def _decompress(f):
        buf.copy_(f.value()[0])
        return buf

    return dist.all_reduce(small, op=dist.ReduceOp.SUM, group=group, async_op=True).get_future().then(_decompress)


# --------------------------------------------------------------------------------------------------------------------

class ArenaDataParallel(torch.nn.Module):
    """``DistributedDataParallel`` contract over a :class:`FlatGradArena` (see the module docstring).

    ``module``: the model (parameters fp32 on one device).  ``arena``: reuse an existing gradient arena (e.g. the one a
    ``FusedAdamW`` was built on) — it must cover exactly the module's trainable parameters.  ``bucket_cap_mb`` /
    ``first_bucket_mb``: segment sizes.  ``reduce_op``: ``"mean"`` (DDP) or ``"sum"`` (fold 1/world into the optimizer).
    ``overlap=False`` launches every segment from the end-of-backward callback instead (one after the other, same result).
    Unused parameters are ALWAYS tolerated (their segments are exchanged by the end-of-backward callback).  What
    ``find_unused_parameters=True`` adds is what it adds in DDP: ``forward`` walks the autograd graph of its output and marks
    the parameters it cannot reach as ready, so a segment that holds a dead deep-supervision head does not hold back the
    segments behind it (launch order is fixed); ``static_graph=True`` does that walk on the first iteration only.
    ``gradient_as_bucket_view`` / ``device_ids`` / ``output_device`` are accepted for signature compatibility (gradients
    always are bucket views)."""

    def __init__(self, module: torch.nn.Module, *, process_group=None, arena: Optional[FlatGradArena] = None,
                 bucket_cap_mb: float = 25.0, first_bucket_mb: Optional[float] = 1.0, reduce_op: str = "mean",
                 overlap: bool = True, broadcast_buffers: bool = True, init_sync: bool = True,
                 find_unused_parameters: bool = True, gradient_as_bucket_view: bool = True, static_graph: bool = False,
                 device_ids=None, output_device=None) -> None:
        super().__init__()
        del gradient_as_bucket_view, device_ids, output_device
        self.find_unused_parameters = bool(find_unused_parameters)
        self.static_graph = bool(static_graph)
        self._static_unused: Optional[List[int]] = None
        if reduce_op not in ("mean", "sum"):
            raise ValueError(f"reduce_op must be 'mean' or 'sum', got {reduce_op!r}")
        self.module = module
        self.process_group = process_group
        self.reduce_op = reduce_op
        self.overlap = bool(overlap)
        self.broadcast_buffers = bool(broadcast_buffers)
        self.require_backward_grad_sync = True
        if init_sync:
            broadcast_parameters(module, src=0, group=process_group)
        trainable = [p for p in module.parameters() if p.requires_grad]
        if arena is None:
            arena = FlatGradArena(trainable)
        elif sorted(id(p) for p in arena.params) != sorted(id(p) for p in trainable):
            raise ValueError("ArenaDataParallel: the gradient arena must cover exactly the module's trainable parameters")
        self.arena = arena
        per_mb = (1 << 20) // 4
        plan = plan_segments([p.numel() for p in arena.params], arena.offsets, arena.total, int(bucket_cap_mb * per_mb),
                             None if first_bucket_mb is None else int(min(first_bucket_mb, bucket_cap_mb) * per_mb))
        self.segments: List[GradSegment] = [GradSegment(arena, k, a, b, lo, hi, k == len(plan) - 1)
                                            for k, (a, b, lo, hi) in enumerate(plan)]
        self._seg_of = [0] * len(arena.params)
        for s in self.segments:
            for i in range(s.first, s.last):
                self._seg_of[i] = s.index()
        self._hook_state: Any = process_group
        self._hook: Callable = allreduce_mean_hook if reduce_op == "mean" else allreduce_sum_hook
        self._handles = [p.register_post_accumulate_grad_hook(self._make_param_hook(i)) for i, p in enumerate(arena.params)]
        self.launch_log: List[tuple] = []          # (segment index, parameters ready when it was launched) of the last step
        self._reset_step()

    # ---- DDP surface -------------------------------------------------------------------------------------------------
    def forward(self, *args, **kwargs):
        armed = self.require_backward_grad_sync and torch.is_grad_enabled()
        if self.broadcast_buffers and armed and self.training:
            self._sync_buffers()
        out = self.module(*args, **kwargs)
        if armed and self.overlap and self.find_unused_parameters:
            self._mark_unreachable(out)
        return out

    def _mark_unreachable(self, out) -> None:
        """DDP's ``prepare_for_backward``: parameters the autograd graph of ``out`` does not reach will get no gradient this
        iteration — count them as ready now (nothing is launched here; the first gradient hook of the backward pass does)."""
        if self.static_graph and self._static_unused is not None:
            unused = self._static_unused
        else:
            tensors, stack = [], [out]
            while stack:
                o = stack.pop()
                if isinstance(o, torch.Tensor):
                    tensors.append(o)
                elif isinstance(o, dict):
                    stack.extend(o.values())
                elif isinstance(o, (list, tuple)):
                    stack.extend(o)
            reached, seen = set(), set()
            nodes = [t.grad_fn for t in tensors if t.grad_fn is not None]
            while nodes:
                fn = nodes.pop()
                if fn in seen:
                    continue
                seen.add(fn)
                var = getattr(fn, "variable", None)            # AccumulateGrad
                if var is not None:
                    reached.add(id(var))
                for nxt, _ in fn.next_functions:
                    if nxt is not None and nxt not in seen:
                        nodes.append(nxt)
            unused = [i for i, p in enumerate(self.arena.params) if id(p) not in reached]
            if self.static_graph:
                self._static_unused = unused
        for i in unused:
            if not self._ready[i]:
                self._attach_view(i)          # BEFORE the segment can be launched: a detached .grad means a stale slice
                self._ready[i] = True
                self._n_ready += 1
                self._missing[self._seg_of[i]] -= 1

    def _attach_view(self, i: int) -> None:
        """Make ``params[i].grad`` the arena view again.  ``zero_grad(set_to_none=True)`` (Lightning's default) leaves
        ``None`` — and does NOT clear the arena, so the slice still holds the previous step's gradient: zero it; a foreign
        tensor (autograd allocated one because ``.grad`` was ``None``) is copied in."""
        p, view = self.arena.params[i], self.arena.view_of(i)
        if p.grad is None:
            view.zero_()
            p.grad = view
        elif p.grad.data_ptr() != view.data_ptr():
            view.copy_(p.grad)
            p.grad = view

    @contextlib.contextmanager
    def no_sync(self):
        """Gradients accumulate locally (into the arena) without an exchange — the non-boundary micro-batches of
        ``accumulate_grad_batches``.  The first backward outside the context exchanges the accumulated sum."""
        old = self.require_backward_grad_sync
        self.require_backward_grad_sync = False
        try:
            yield
        finally:
            self.require_backward_grad_sync = old

    def register_comm_hook(self, state: Any, hook: Callable) -> None:
        """``hook(state, bucket) -> Future[Tensor]``; the future's tensor is written back into the bucket when it is a
        different tensor (compression hooks).  With a custom hook ``reduce_op`` no longer applies: the hook owns the scale."""
        if not callable(hook):
            raise TypeError("Communication hook must be callable.")
        self._hook_state, self._hook = state, hook

    def zero_grad(self, set_to_none: bool = False) -> None:      # keeps the views (Lightning passes set_to_none=True)
        self.arena.zero_grad(set_to_none)

    def remove_hooks(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []

    # ---- the exchange ------------------------------------------------------------------------------------------------
    def _sync_buffers(self) -> None:
        if _world(self.process_group) == 1:
            return
        with torch.no_grad():
            for b in self.module.buffers():
                if b.is_floating_point():
                    dist.broadcast(b.data, src=0, group=self.process_group)

    def _reset_step(self) -> None:
        self._ready = [False] * len(self.arena.params)
        self._missing = [s.last - s.first for s in self.segments]
        self._next = 0
        self._n_ready = 0
        self._futures: List[tuple] = []
        self._queued = False

    def _make_param_hook(self, i: int) -> Callable:
        def hook(p: torch.Tensor) -> None:
            if not self.require_backward_grad_sync:
                return
            self._attach_view(i)                                                 # a detached view (zero_grad(set_to_none))
            if not self._queued:
                self._queued = True
                self.launch_log = []
                torch.autograd.Variable._execution_engine.queue_callback(self._finalize)
            if not self._ready[i]:
                self._ready[i] = True
                self._n_ready += 1
                self._missing[self._seg_of[i]] -= 1
            elif self._seg_of[i] < self._next:
                raise RuntimeError(f"ArenaDataParallel: parameter {i} received a gradient after its segment was exchanged "
                                   "(static_graph=True with a graph that changed, or two backward passes per forward)")
            if self.overlap:
                self._launch_ready()
        return hook

    def _launch(self, seg: GradSegment) -> None:
        self.launch_log.append((seg.index(), self._n_ready))
        if self._hook in (allreduce_mean_hook, allreduce_sum_hook) and _world(self.process_group) == 1:
            return                                                               # single process: nothing to exchange
        fut = self._hook(self._hook_state, seg)
        self._futures.append((seg, fut))

    def _launch_ready(self) -> None:
        # strictly in segment order: every rank issues the same sequence of collectives whatever its gradients' order
        while self._next < len(self.segments) and self._missing[self._next] == 0:
            self._launch(self.segments[self._next])
            self._next += 1

    def _finalize(self) -> None:
        try:
            # parameters neither hooked nor marked unused belong to segments that are not launched yet (a segment goes out
            # only when all of its parameters are ready): repair their views now, never those of an exchanged segment
            for i in range(len(self.arena.params)):
                if not self._ready[i]:
                    self._attach_view(i)
            while self._next < len(self.segments):
                self._launch(self.segments[self._next])
                self._next += 1
            for seg, fut in self._futures:
                fut.wait()                                      # CUDA futures: orders the current stream, no host block
                out = fut.value()
                if isinstance(out, (list, tuple)):
                    out = out[0]
                seg.set_buffer(out)
        finally:
            self._reset_step()

    def reduce_now(self) -> None:
        """Exchange outside of a backward pass (e.g. gradients written by a CUDA-graph replay, which runs no hooks)."""
        self.launch_log = []
        for i in range(len(self.arena.params)):
            if not (self._ready[i] and self._seg_of[i] < self._next):
                self._attach_view(i)
        self._ready = [True] * len(self.arena.params)
        self._n_ready = len(self._ready)
        self._finalize()


def make_arena_ddp_strategy(strategy_kwargs: Optional[dict] = None, **wrapper_kwargs):
    """A Lightning ``DDPStrategy`` whose wrapped model is :class:`ArenaDataParallel` (the "thin custom Strategy" of SURVEY
    §8(b)3): pass it as ``pl.Trainer(strategy=...)`` where ``trainer.py:241-249`` builds ``DDPStrategy(find_unused_parameters=
    ...)``; ``strategy_kwargs`` go to ``DDPStrategy.__init__``, the rest to :class:`ArenaDataParallel`.  Lightning is not part of this image, so the class is created on demand and raises ``ImportError`` without it."""
    try:
        from lightning.pytorch.strategies import DDPStrategy  # type: ignore
    except ImportError:                                       # the reference imports ``pytorch_lightning``
        from pytorch_lightning.strategies import DDPStrategy  # type: ignore

    class ArenaDDPStrategy(DDPStrategy):                      # pragma: no cover - needs Lightning
        strategy_name = "pcb200_arena_ddp"

        def _setup_model(self, model):
            kw = {k: v for k, v in getattr(self, "_ddp_kwargs", {}).items() if k in ("bucket_cap_mb", "broadcast_buffers")}
            kw.update(wrapper_kwargs)
            return ArenaDataParallel(model, process_group=getattr(self, "_process_group", None), **kw)

        def _register_ddp_hooks(self) -> None:                # comm hooks go through ArenaDataParallel.register_comm_hook
            hook = getattr(self, "_ddp_comm_hook", None)
            if hook is not None:
                self.model.register_comm_hook(getattr(self, "_ddp_comm_state", None), hook)

        @contextlib.contextmanager
        def block_backward_sync(self):
            if isinstance(self.model, ArenaDataParallel):
                with self.model.no_sync():
                    yield None
            else:
                yield None

    return ArenaDDPStrategy(**dict(strategy_kwargs or {}))
