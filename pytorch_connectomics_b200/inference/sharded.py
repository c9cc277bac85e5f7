"""Multi-GPU sliding-window inference of ONE volume: contiguous z-slab ownership of the eager window grid
with a neighbour exchange of the overlap planes.

The reference shards a single volume only on its lazy path, by interleaving windows over ranks and
reducing two FULL-volume accumulators onto rank 0 (``connectomics/inference/lazy.py:1077-1104``,
``lazy_distributed.py:35-169``: O(world x volume) traffic), or by halo-extended chunks that recompute the
windows straddling a chunk face (``inference/chunked.py:437-723``).  Here every window of the eager grid
(``window.py:92-134``) is computed exactly once:

* the distinct z-starts of the grid are split into ``world`` contiguous groups; rank r runs the windows whose
  z-start is in its group and accumulates ``value += pred*w, weight += w`` into a LOCAL slab that only spans
  its windows' z-extent (``[first z-start, last z-start + roi_z)``);
* output plane ``z`` is owned by the rank whose group starts at or before it (``own`` ranges partition
  ``[0, D)``); the planes of a slab that another rank owns (``roi_z - stride_z`` planes per face, 80 of 160 at
  50 % overlap) are sent to that owner — one ``send/recv`` pair per face over NCCL/NVLink, value and weight
  packed in one message — and added there;
* each rank normalises and returns its own z-range (``value / clamp_min(weight, 1e-4)``, ``window.py:275-294``).

Integer planning is pure host logic (``plan_z_slabs``) and bit-exact by construction: the union of the per-rank
window lists IS the eager grid.  Floating-point sums in the exchanged planes associate as
(own windows) + (neighbour's windows) instead of strictly in grid order; everywhere else, and for world == 1,
the result is bit-identical to ``EagerSlidingWindowEngine``.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .. import _lib as L
from . import window as W


@dataclass
class SlabPlan:
    """What one rank does.  All z coordinates are in the (grown) volume; ``hi`` is exclusive."""
    rank: int
    world: int
    image: Tuple[int, int, int]
    roi: Tuple[int, int, int]
    z_starts: List[int] = field(default_factory=list)
    windows: List[Tuple[int, int, int]] = field(default_factory=list)   # eager-grid starts owned by this rank
    slab: Tuple[int, int] = (0, 0)            # z-range the local accumulators cover
    own: Tuple[int, int] = (0, 0)             # z-range of the output this rank normalises and returns
    sends: List[Tuple[int, int, int]] = field(default_factory=list)     # (peer, z_lo, z_hi): my partial planes -> peer
    recvs: List[Tuple[int, int, int]] = field(default_factory=list)     # (peer, z_lo, z_hi): peer's partial planes -> me


def split_contiguous(n_items: int, parts: int) -> List[Tuple[int, int]]:
    """[lo, hi) index ranges of ``parts`` contiguous groups, sizes differing by at most one, larger groups first
    (25 z-starts over 8 ranks -> 4,3,3,3,3,3,3,3)."""
    base, extra = divmod(n_items, parts)
    out, lo = [], 0
    for p in range(parts):
        n = base + (1 if p < extra else 0)
        out.append((lo, lo + n))
        lo += n
    return out


def plan_z_slabs(image_size: Sequence[int], roi_size: Sequence[int], overlap, world: int) -> List[SlabPlan]:
    """Plans for every rank.  ``image_size`` must already be grown to at least ``roi_size``."""
    image = tuple(int(v) for v in image_size)
    roi = tuple(int(v) for v in roi_size)
    if len(image) != 3 or len(roi) != 3:
        raise ValueError(f"z-slab sharding is defined for 3-D volumes; got image {image}, roi {roi}")
    if world < 1:
        raise ValueError(f"world must be >= 1, got {world}")
    if any(i < r for i, r in zip(image, roi)):
        raise ValueError(f"image {image} must be grown to the roi {roi} before planning")
    starts = W._plan(L.GRID_EAGER, image, roi, overlap)
    zs = sorted({s[0] for s in starts})
    groups = split_contiguous(len(zs), world)
    plans = [SlabPlan(rank=r, world=world, image=image, roi=roi) for r in range(world)]
    live = []
    for r, (lo, hi) in enumerate(groups):
        if hi > lo:
            mine = zs[lo:hi]
            p = plans[r]
            p.z_starts = mine
            first, last = mine[0], mine[-1]
            p.windows = [s for s in starts if first <= s[0] <= last]   # grid order preserved (z-major)
            p.slab = (first, last + roi[0])
            live.append(r)
    for i, r in enumerate(live):
        p = plans[r]
        own_lo = 0 if i == 0 else p.slab[0]
        own_hi = image[0] if i == len(live) - 1 else plans[live[i + 1]].slab[0]
        p.own = (own_lo, own_hi)
    for r in live:
        for q in live:
            if q == r:
                continue
            lo, hi = max(plans[r].slab[0], plans[q].own[0]), min(plans[r].slab[1], plans[q].own[1])
            if hi > lo:
                plans[r].sends.append((q, lo, hi))
                plans[q].recvs.append((r, lo, hi))
    return plans


def _pack(value: torch.Tensor, weight: torch.Tensor, lo: int, hi: int) -> torch.Tensor:
    """[Cout+1, hi-lo, H, W] contiguous message: value planes then the weight planes."""
    return torch.cat([value[0, :, lo:hi], weight[0, :, lo:hi]], dim=0).contiguous()


def exchange_overlaps(value: torch.Tensor, weight: torch.Tensor, plan: SlabPlan,
                      group: Optional[dist.ProcessGroup] = None, comm=None) -> None:
    """Send the planes of my slab that another rank owns and add what others computed for my planes (in place).
    Works on any backend (NCCL on GPUs, gloo in the CPU tests); every rank of ``group`` must call it.  ``comm``: a
    :class:`pytorch_connectomics_b200.comm.NativeComm` — the same messages as ONE grouped send/recv through the C ABI
    (``pcb_sw_exchange_overlap``) instead of ``torch.distributed``; ranks in ``plan`` are ranks of that communicator."""
    if not plan.sends and not plan.recvs:
        return
    z0 = plan.slab[0]
    cout = int(value.shape[1])
    ops, inbox = [], []
    if comm is not None:
        outbox = [(_pack(value, weight, lo - z0, hi - z0), peer) for peer, lo, hi in plan.sends]
        for peer, lo, hi in plan.recvs:
            inbox.append((torch.empty((cout + 1, hi - lo, *value.shape[3:]), device=value.device, dtype=value.dtype), lo, hi))
        comm.exchange(outbox, [(b, peer) for (b, _, _), (peer, _, _) in zip(inbox, plan.recvs)])
    else:
        for peer, lo, hi in plan.sends:
            ops.append(dist.P2POp(dist.isend, _pack(value, weight, lo - z0, hi - z0), peer, group))
        for peer, lo, hi in plan.recvs:
            buf = torch.empty((cout + 1, hi - lo, *value.shape[3:]), device=value.device, dtype=value.dtype)
            inbox.append((buf, lo, hi))
            ops.append(dist.P2POp(dist.irecv, buf, peer, group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for buf, lo, hi in inbox:      # fixed (rank-ordered) accumulation order -> deterministic
        value[0, :, lo - z0:hi - z0] += buf[:cout]
        weight[0, :, lo - z0:hi - z0] += buf[cout:]


class ZSlabShardedEngine:
    """``engine(inputs=[1,C,D,H,W], network=fn) -> ([1,Cout,own_hi-own_lo,H,W], (own_lo, own_hi))`` on every rank.

    Same constructor arguments as ``EagerSlidingWindowEngine`` (``window.py:530-561``) plus the process group.
    ``inputs`` may live on the host: only this rank's slab is copied to the GPU."""

    def __init__(self, *, roi_size, sw_batch_size: int, overlap, mode: str, padding_mode: str = "constant",
                 cval: float = 0.0, device=None, group: Optional[dist.ProcessGroup] = None,
                 rank: Optional[int] = None, world: Optional[int] = None, cuda_graph: bool = True, comm=None) -> None:
        self.roi = tuple(int(v) for v in roi_size)
        if len(self.roi) != 3:
            raise ValueError(f"ZSlabShardedEngine needs a 3-D roi_size, got {roi_size}")
        self.sw_batch_size = max(1, int(sw_batch_size))
        self.overlap = overlap
        self.mode = W._normalize_blending_mode(mode)
        self.padding_mode = padding_mode
        self.cval = float(cval)
        self.device = device
        self.group = group
        self.comm = comm                      # NativeComm: exchange through the C ABI instead of torch.distributed
        self.cuda_graph = bool(cuda_graph)
        if comm is not None:
            rank = comm.rank if rank is None else rank
            world = comm.world if world is None else world
        use_dist = dist.is_available() and dist.is_initialized()
        self.rank = rank if rank is not None else (dist.get_rank(group) if use_dist else 0)
        self.world = world if world is not None else (dist.get_world_size(group) if use_dist else 1)

    # ---- phase 1: my windows -> local slab accumulators
    def accumulate_local(self, inputs: torch.Tensor, network: Callable[[torch.Tensor], torch.Tensor], plan: SlabPlan,
                         presliced: bool = False):
        """``inputs`` is the whole volume (only this rank's z-range is copied to the GPU) or, with ``presliced``, the
        slab ``[1, C, plan.slab extent, H, W]`` itself (host or device)."""
        roi = self.roi
        dev = W._device_or_raise(self.device if self.device is not None else inputs.device)
        z0, z1 = plan.slab
        if presliced:
            if int(inputs.shape[2]) != z1 - z0:
                raise ValueError(f"ZSlabShardedEngine: slab has {int(inputs.shape[2])} planes, the plan needs {z1 - z0}")
            vol = inputs.to(dev, non_blocking=True)
        else:
            vol = inputs[:, :, z0:z1].to(dev, non_blocking=True)
        local_image = (z1 - z0, plan.image[1], plan.image[2])
        starts = [(s[0] - z0, s[1], s[2]) for s in plan.windows]
        value, weight, _ = W.run_window_list(vol, network, starts, roi=roi, image=local_image,
                                             sw_batch_size=self.sw_batch_size, padding_mode=self.padding_mode,
                                             cval=self.cval, mode=self.mode, sw_device=dev, work_device=dev,
                                             probe_first=False, cuda_graph=self.cuda_graph, who="ZSlabShardedEngine")
        return value, weight

    def run_slab(self, slab: torch.Tensor, network: Callable[[torch.Tensor], torch.Tensor], plan: SlabPlan):
        """One rank's share given only ITS slab of the volume (what a loader that reads per-rank z-ranges hands over):
        windows -> neighbour exchange -> normalised own planes (``None`` for a rank without windows)."""
        if not plan.windows:
            if self.world > 1:
                exchange_overlaps(torch.empty(0), torch.empty(0), plan, self.group, self.comm)
            return None
        value, weight = self.accumulate_local(slab, network, plan, presliced=True)
        if self.world > 1:
            exchange_overlaps(value, weight, plan, self.group, self.comm)
        return self.finalize(value, weight, plan)

    # ---- phase 3: normalise my own planes
    @staticmethod
    def finalize(value: torch.Tensor, weight: torch.Tensor, plan: SlabPlan) -> torch.Tensor:
        z0 = plan.slab[0]
        lo, hi = plan.own[0] - z0, plan.own[1] - z0
        v = value[:, :, lo:hi].contiguous()
        w = weight[:, :, lo:hi].contiguous()
        return W.normalize_weighted_accumulator(v, w)

    def __call__(self, inputs: torch.Tensor, network: Callable[[torch.Tensor], torch.Tensor]):
        if inputs.dim() != 5:
            raise ValueError("ZSlabShardedEngine: inputs must have shape (1, C, D, H, W); "
                             f"got shape {tuple(inputs.shape)}.")
        if inputs.shape[0] != 1:
            raise ValueError(f"ZSlabShardedEngine expects batch size 1; got batch {inputs.shape[0]}.")
        image = tuple(int(v) for v in inputs.shape[2:])
        if any(i < r for i, r in zip(image, self.roi)):
            raise ValueError(f"ZSlabShardedEngine: volume {image} is smaller than the window {self.roi}; "
                             "use EagerSlidingWindowEngine (it pads up to the window) for such inputs.")
        plan = plan_z_slabs(image, self.roi, self.overlap, self.world)[self.rank]
        if not plan.windows:       # more ranks than z-starts: nothing to do here, but stay in the collective
            if self.world > 1:
                exchange_overlaps(torch.empty(0), torch.empty(0), plan, self.group, self.comm)
            return None, plan.own
        value, weight = self.accumulate_local(inputs, network, plan)
        if self.world > 1:
            exchange_overlaps(value, weight, plan, self.group, self.comm)
        return self.finalize(value, weight, plan), plan.own


__all__ = ["SlabPlan", "ZSlabShardedEngine", "exchange_overlaps", "plan_z_slabs", "split_contiguous"]
