"""Distributed reduction of sliding-window accumulators (``connectomics/inference/lazy_distributed.py``).

The reference stages full-volume CPU accumulators through the GPU in 128 MB chunks and ``reduce``s them
to rank 0 (``:78-107``).  Here the accumulators already live in HBM, so the exchange is one NCCL
``reduce(SUM, dst=0)`` per accumulator over NVLink (no host staging); non-root ranks get ``None`` back,
which makes ``lazy_sliding_window(..., accumulator_reduce=hook)`` return an empty tensor exactly like the
reference's non-root path (``lazy.py:1241-1249``).  The sanity checks of ``:42-75,110-129`` (all ranks
agree on the accumulator shape, no rank has an empty shard) are kept as all-gathers of a few int64s.
"""

from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def should_shard_windows(enable: bool) -> bool:
    """``lazy_distributed.py:16-31`` — window sharding is on only inside an initialised multi-rank group."""
    return bool(enable) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def validate_patch_shard(local_count: int, total_count: int, device) -> None:
    """``lazy_distributed.py:110-129`` — every rank must own at least one window."""
    world = dist.get_world_size()
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([int(local_count)], dtype=torch.int64, device=device))
    empty = [r for r, c in enumerate(counts) if int(c.item()) == 0]
    if empty:
        raise RuntimeError(f"Distributed lazy sliding-window sharding produced empty shards on ranks {empty} "
                           f"({total_count} windows over {world} ranks).")


def make_accumulator_reducer(group: Optional[dist.ProcessGroup] = None
                             ) -> Callable[[torch.Tensor, torch.Tensor], Optional[Tuple[torch.Tensor, torch.Tensor]]]:
    """Returns the ``accumulator_reduce`` hook: SUM both accumulators onto rank 0 (``:132-169``)."""

    def reduce(value: torch.Tensor, weight: torch.Tensor):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return value, weight
        world = dist.get_world_size(group)
        shape = torch.tensor(list(value.shape) + list(weight.shape), dtype=torch.int64, device=value.device)
        shapes = [torch.zeros_like(shape) for _ in range(world)]
        dist.all_gather(shapes, shape, group=group)
        if any(not torch.equal(s, shape) for s in shapes):
            raise RuntimeError("Distributed lazy sliding-window ranks disagree on the accumulator shape: "
                               f"{[s.tolist() for s in shapes]}")
        dist.reduce(value, dst=0, op=dist.ReduceOp.SUM, group=group)
        dist.reduce(weight, dst=0, op=dist.ReduceOp.SUM, group=group)
        return (value, weight) if dist.get_rank(group) == 0 else None

    return reduce


__all__ = ["should_shard_windows", "validate_patch_shard", "make_accumulator_reducer"]
