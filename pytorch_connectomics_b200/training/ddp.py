"""DDP seam on the B200 box: one flat gradient arena + ONE NCCL all-reduce per optimizer step.

The reference gets data parallelism from Lightning's ``DDPStrategy`` -> ``torch.nn.parallel.
DistributedDataParallel`` (``connectomics/training/lightning/trainer.py:231-256,314-334``): bucketed
all-reduce of every gradient, ``find_unused_parameters=True`` for MedNeXt (unused deep-supervision
heads and ``dummy_tensor``).  Here every parameter's ``.grad`` is a view into one contiguous fp32 arena,
so autograd accumulates straight into it (in place), parameters that receive no gradient simply stay
zero (the ``find_unused_parameters`` semantics without the graph walk), and the exchange step is a single
``all_reduce(SUM)`` over NVLink followed by a scale by 1/world (gradient mean, as DDP).  With
``accumulate_grad_batches`` > 1 call :func:`allreduce_gradients` only on the boundary micro-step.
"""

from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class FlatGradArena:
    """``align``: every parameter's slice starts at a multiple of ``align`` elements (default 64 = 256 B), so slices of
    the arena — and of the parameter / optimizer-state arenas laid out the same way (``training/optim.py``) — can be
    handed to kernels that use 128-bit accesses; padding elements stay zero."""

    def __init__(self, params: Iterable[torch.nn.Parameter], align: int = 64):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradArena: no trainable parameters")
        dev = self.params[0].device
        self.align = max(1, int(align))
        self.offsets: List[int] = []
        off = 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("FlatGradArena expects fp32 parameters on one device")
            self.offsets.append(off)
            off += -(-p.numel() // self.align) * self.align
        self.total = off
        self.buffer = torch.zeros(self.total, device=dev, dtype=torch.float32)
        for p, o in zip(self.params, self.offsets):
            p.grad = self.buffer[o:o + p.numel()].view_as(p)

    def view_of(self, index: int, flat: Optional[torch.Tensor] = None) -> torch.Tensor:
        """the slice of ``flat`` (default: the gradient arena) that belongs to parameter ``index``"""
        p, o = self.params[index], self.offsets[index]
        return (self.buffer if flat is None else flat)[o:o + p.numel()].view_as(p)

    def packed(self) -> torch.Tensor:
        """the gradients concatenated without padding (parameter order)"""
        return torch.cat([self.view_of(i).reshape(-1) for i in range(len(self.params))])

    def zero(self) -> None:
        """Use instead of ``optimizer.zero_grad()`` (which would detach the views when set_to_none)."""
        self.buffer.zero_()

    def zero_grad(self, set_to_none: bool = False) -> None:
        """``optimizer.zero_grad`` / ``model.zero_grad`` shim: zeroes the arena and keeps every ``.grad`` a view of it
        (``set_to_none`` is accepted and ignored — detaching the views is exactly what must not happen)."""
        del set_to_none
        self.rebind()
        self.buffer.zero_()

    def patch_zero_grad(self, optimizer: torch.optim.Optimizer) -> None:
        """Route ``optimizer.zero_grad()`` (Lightning calls it every step with ``set_to_none=True``) to the arena."""
        optimizer.zero_grad = self.zero_grad  # type: ignore[method-assign]

    def gather_stray_grads(self) -> int:
        """If something replaced a ``.grad`` view (``zero_grad(set_to_none=True)`` followed by a backward pass leaves
        autograd-allocated tensors), copy those gradients into the arena and re-attach the views, so the exchange step
        never reduces a stale buffer.  Returns the number of repaired parameters."""
        fixed = 0
        for i, p in enumerate(self.params):
            view = self.view_of(i)
            g = p.grad
            if g is None:
                view.zero_()
                p.grad = view
                fixed += 1
            elif g.data_ptr() != view.data_ptr():
                view.copy_(g)
                p.grad = view
                fixed += 1
        return fixed

    def rebind(self) -> None:
        """Re-attach ``.grad`` views if something replaced them (e.g. ``zero_grad(set_to_none=True)``)."""
        for i, p in enumerate(self.params):
            view = self.view_of(i)
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                p.grad = view

    @staticmethod
    def _world(group) -> int:
        return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1

    def allreduce_sum(self, group: Optional[dist.ProcessGroup] = None, comm=None) -> None:
        """The exchange step alone (NCCL all-reduce SUM over NVLink); pair with :meth:`scale_mean`.  ``comm``: a
        :class:`pytorch_connectomics_b200.comm.NativeComm` — the same exchange through the C ABI (``pcb_grad_allreduce``)
        instead of ``torch.distributed``."""
        if comm is not None:
            if comm.world > 1:
                comm.allreduce_(self.buffer, 1.0)
            return
        if self._world(group) > 1:
            dist.all_reduce(self.buffer, op=dist.ReduceOp.SUM, group=group)

    def scale_mean(self, group: Optional[dist.ProcessGroup] = None) -> None:
        if self._world(group) > 1:
            self.buffer.mul_(1.0 / self._world(group))

    def allreduce(self, group: Optional[dist.ProcessGroup] = None, comm=None) -> None:
        self.gather_stray_grads()       # host-side pointer checks; copies only when a view was replaced
        if comm is not None:            # SUM and 1/world in one library call
            comm.allreduce_(self.buffer, 1.0 / comm.world)
            return
        self.allreduce_sum(group)
        self.scale_mean(group)


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group: Optional[dist.ProcessGroup] = None) -> None:
    """What ``DistributedDataParallel`` does at construction (``trainer.py:231-256`` builds a ``DDPStrategy``): every
    rank starts from rank ``src``'s parameters AND buffers (BatchNorm running statistics of ``monai_unet``), so replicas
    that average gradients are replicas of the same model.  No-op without an initialised process group / world 1."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            if t.is_floating_point() or t.dtype in (torch.int64, torch.int32):
                dist.broadcast(t.data, src=src, group=group)


def allreduce_gradients(arena: FlatGradArena, group: Optional[dist.ProcessGroup] = None) -> None:
    arena.allreduce(group)
