"""Distributed reduction of sliding-window accumulators — the public names and signatures of
``connectomics/inference/lazy_distributed.py`` (``distributed_context``, ``is_distributed_window_sharding_enabled``,
``distributed_reduction_device``, ``validate_distributed_tensor_shape``, ``reduce_cpu_tensor_to_rank_zero``,
``validate_distributed_patch_shard``, ``make_accumulator_reduce_hook``), so ``lazy.py`` / ``tta.py`` callers import them unchanged.

The reference keeps full-volume accumulators on the HOST and stages them through the GPU in ``chunk_mb`` pieces for a
``reduce`` to rank 0 (``:78-107``).  The engine here accumulates in HBM, so an accumulator that already sits on
``reduction_device`` is reduced in place with ONE collective over NVLink and no staging; a host accumulator still takes the
reference's chunked route (that is also what the gloo tests exercise on the CPU).  Non-root ranks get ``None`` back, which
makes the lazy engine return an empty tensor exactly like the reference's non-root path (``lazy.py:1241-1249``).  The sanity
checks (all ranks agree on the accumulator shape ``:42-75``, no rank has an empty shard ``:110-129``) are all-gathers of a few
int64s with the reference's error texts.
"""

from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist

_MAX_RANK = 8       # tensor rank the shape check can describe (lazy_distributed.py:52)


def distributed_context() -> Tuple[bool, int, int]:
    """(is_distributed, rank, world_size) of the default group; ``(False, 0, 1)`` outside one (``:10-13``)."""
    if dist.is_available() and dist.is_initialized():
        return True, dist.get_rank(), dist.get_world_size()
    return False, 0, 1


def is_distributed_window_sharding_enabled(cfg) -> bool:
    """``:16-31`` — on only for a lazy (zarr / h5) loader with ``inference.sliding_window.distributed_sharding`` inside an
    initialised group of more than one rank."""
    sliding = getattr(getattr(cfg, "inference", None), "sliding_window", None)
    if sliding is None or not getattr(sliding, "distributed_sharding", False):
        return False
    loader = getattr(getattr(cfg, "data", None), "dataloader", None)
    lazy = bool(getattr(loader, "use_lazy_zarr", False) or getattr(loader, "use_lazy_h5", False))
    active, _, world = distributed_context()
    return bool(lazy and active and world > 1)


def distributed_reduction_device(infer_device: torch.device) -> torch.device:
    """``:34-39`` — collectives run on the inference GPU, else the current GPU, else the CPU (gloo)."""
    infer_device = torch.device(infer_device)
    if infer_device.type == "cuda":
        return infer_device
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def _gather_int64(values: List[int], device, group=None) -> List[List[int]]:
    world = dist.get_world_size(group)
    mine = torch.tensor(values, dtype=torch.int64, device=device)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [[int(v) for v in t.tolist()] for t in out]


def validate_distributed_tensor_shape(tensor: torch.Tensor, *, name: str, reduction_device: torch.device, group=None) -> None:
    """``:42-75`` — every rank must reduce a tensor of the same shape."""
    active, _, _ = distributed_context()
    if not active:
        return
    if tensor.ndim > _MAX_RANK:
        raise RuntimeError(f"{name} has rank {tensor.ndim}, exceeding supported rank {_MAX_RANK}.")
    record = [int(tensor.ndim)] + [int(d) for d in tensor.shape] + [-1] * (_MAX_RANK - tensor.ndim)
    shapes = [tuple(r[1:1 + r[0]]) for r in _gather_int64(record, reduction_device, group)]
    if any(s != shapes[0] for s in shapes[1:]):
        summary = ", ".join(f"rank {r}: {s}" for r, s in enumerate(shapes))
        raise RuntimeError(f"Distributed lazy sliding-window sharding requires every rank to reduce {name} "
                           f"with the same shape, got {summary}.")


def reduce_cpu_tensor_to_rank_zero(tensor: torch.Tensor, *, op, reduction_device: torch.device, chunk_mb: int, name: str,
                                   group=None) -> Optional[torch.Tensor]:
    """``:78-107`` — reduce an accumulator onto rank 0; returns the reduced tensor there and ``None`` elsewhere (the input
    itself outside a process group).  An accumulator that already lives on ``reduction_device`` (the B200 engine's HBM
    accumulators) is reduced in place by one collective; anything else goes through ``reduction_device`` in ``chunk_mb``
    pieces like the reference."""
    active, _, _ = distributed_context()
    if not active:
        return tensor
    rank = dist.get_rank(group)
    reduction_device = torch.device(reduction_device)
    validate_distributed_tensor_shape(tensor, name=name, reduction_device=reduction_device, group=group)
    same = tensor.device.type == reduction_device.type and (reduction_device.index is None or tensor.device == reduction_device)
    if same and tensor.is_contiguous():
        dist.reduce(tensor, dst=dist.get_global_rank(group, 0) if group is not None else 0, op=op, group=group)
        return tensor if rank == 0 else None
    flat = tensor.contiguous().view(-1)
    step = max(1, (max(1, int(chunk_mb or 128)) << 20) // max(1, flat.element_size()))
    result = torch.empty_like(flat) if rank == 0 else None
    dst = dist.get_global_rank(group, 0) if group is not None else 0
    for lo in range(0, flat.numel(), step):
        piece = flat[lo:lo + step].to(device=reduction_device)
        dist.reduce(piece, dst=dst, op=op, group=group)
        if result is not None:
            result[lo:lo + step].copy_(piece)
    return result.view_as(tensor) if result is not None else None


def validate_distributed_patch_shard(*, local_count: int, total_count: int, reduction_device: torch.device, group=None) -> None:
    """``:110-129`` — every rank must own at least one window."""
    active, _, _ = distributed_context()
    if not active:
        return
    counts = [r[0] for r in _gather_int64([int(local_count)], reduction_device, group)]
    if any(c <= 0 for c in counts):
        raise RuntimeError("Distributed lazy sliding-window sharding assigned an empty window shard "
                           f"(total_windows={total_count}, per_rank={counts}). Use fewer GPUs or a "
                           "smaller inference.sliding_window.window_size.")


def make_accumulator_reduce_hook(*, reduction_device: torch.device, chunk_mb: int, group=None
                                 ) -> Callable[[torch.Tensor, torch.Tensor], Optional[Tuple[torch.Tensor, torch.Tensor]]]:
    """``:132-169`` — ``hook(value, weight) -> (value, weight)`` on rank 0, ``None`` on every other rank."""

    def hook(value: torch.Tensor, weight: torch.Tensor):
        kw = dict(op=dist.ReduceOp.SUM, reduction_device=reduction_device, chunk_mb=chunk_mb, group=group)
        v = reduce_cpu_tensor_to_rank_zero(value, name="value accumulator", **kw)
        w = reduce_cpu_tensor_to_rank_zero(weight, name="weight accumulator", **kw)
        return None if v is None or w is None else (v, w)

    return hook


# ---- the shorter forms the engine modules of this package use -------------------------------------------------------
def should_shard_windows(enable: bool) -> bool:
    """window sharding is on only inside an initialised multi-rank group (the cfg-free core of ``:16-31``)"""
    active, _, world = distributed_context()
    return bool(enable) and active and world > 1


def validate_patch_shard(local_count: int, total_count: int, device) -> None:
    validate_distributed_patch_shard(local_count=local_count, total_count=total_count, reduction_device=torch.device(device))


def make_accumulator_reducer(group: Optional[dist.ProcessGroup] = None, chunk_mb: int = 128):
    """the reduce hook for accumulators wherever they live (reduced on their own device when that is a GPU)"""

    def reduce(value: torch.Tensor, weight: torch.Tensor):
        hook = make_accumulator_reduce_hook(reduction_device=distributed_reduction_device(value.device), chunk_mb=chunk_mb,
                                            group=group)
        return hook(value, weight)

    return reduce


__all__ = ["distributed_context", "distributed_reduction_device", "make_accumulator_reduce_hook",
           "is_distributed_window_sharding_enabled", "reduce_cpu_tensor_to_rank_zero", "validate_distributed_patch_shard",
           "validate_distributed_tensor_shape", "should_shard_windows", "validate_patch_shard", "make_accumulator_reducer"]
