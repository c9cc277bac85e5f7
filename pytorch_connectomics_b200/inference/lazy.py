"""Lazy-grid sliding-window inference over a bounded region — the tile loop of
``connectomics/inference/lazy.py:986-1258`` (``_lazy_sliding_window`` / ``lazy_predict_region`` /
``lazy_predict_volume``) on the B200 engine.

The reference's lazy path reads windows from disk (h5/zarr/tiff accessors — out of scope, they need
h5py/zarr) and blends on the CPU.  Here the volume is a tensor (resident in HBM, or pinned host memory
staged to the GPU once); the grid, the clipped boxes, rank sharding and the accumulate/normalise
arithmetic follow the reference exactly:

  * window offsets with face-centred boundary windows (``lazy.py:269-334``, negative starts) from
    ``pcb_sw_plan(PCB_GRID_LAZY | PCB_GRID_LAZY_SNAP)``, filtered to the windows that intersect the
    requested region (``:337-365``);
  * patches are read with the outer padding mode, cast to fp32, run through ``network`` and cast to
    the output dtype (``:1184-1206``); only the part of each window inside the region is accumulated
    (``:1077-1102,1216-1227``);
  * ``rank``/``world_size`` shard the records as ``records[rank::world_size]`` (``:1104``) and
    ``accumulator_reduce`` sees the un-normalised accumulators (``:1241-1249``).
"""

from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch

from .. import _lib as L
from . import window as W


def lazy_window_records(image_size, roi_size, overlap, region_start, region_stop, snap_to_edge: bool):
    """(patch_start, pred_lo, box, out_lo) per window intersecting the region, in grid order."""
    kind = L.GRID_LAZY_SNAP if snap_to_edge else L.GRID_LAZY
    starts = W._plan(kind, image_size, roi_size, overlap, region=(region_start, region_stop))
    recs = []
    for ps in starts:
        lo = tuple(max(ps[a], int(region_start[a])) for a in range(3))
        hi = tuple(min(ps[a] + int(roi_size[a]), int(region_stop[a])) for a in range(3))
        if any(hi[a] <= lo[a] for a in range(3)):
            continue
        recs.append((tuple(ps), tuple(lo[a] - ps[a] for a in range(3)), tuple(hi[a] - lo[a] for a in range(3)),
                     tuple(lo[a] - int(region_start[a]) for a in range(3))))
    return recs


def lazy_sliding_window(volume: torch.Tensor, network: Callable[[torch.Tensor], torch.Tensor], *, roi_size,
                        overlap=0.5, mode: str = "bump", padding_mode: str = "constant", cval: float = 0.0,
                        region_start: Optional[Sequence[int]] = None, region_stop: Optional[Sequence[int]] = None,
                        snap_to_edge: bool = False, sw_batch_size: int = 1, output_dtype: torch.dtype = torch.float32,
                        border_mask: Sequence[int] = (), rank: int = 0, world_size: int = 1,
                        accumulator_reduce=None, device=None, normalize: bool = True):
    """Returns the blended prediction of the region ``[1, Cout, *region]`` on the compute device
    (or ``(value, weight)`` un-normalised when ``normalize=False``)."""
    roi = tuple(int(v) for v in roi_size)
    if len(roi) != 3:
        raise ValueError(f"Lazy sliding-window inference currently supports 3D only, got {roi}.")
    if volume.dim() != 5 or volume.shape[0] != 1:
        raise ValueError(f"expected a [1, C, D, H, W] volume, got {tuple(volume.shape)}")
    dev = torch.device(device) if device is not None else volume.device
    W._device_or_raise(dev)
    vol = volume.to(dev, non_blocking=True)
    img = tuple(int(v) for v in vol.shape[-3:])
    if any(img[a] < roi[a] for a in range(3)):
        raise ValueError("Lazy sliding-window inference requires the volume to be at least as large as the ROI "
                         f"in every axis. Got bounds_shape={img}, roi_size={roi}.")
    start = (0, 0, 0) if region_start is None else tuple(max(0, int(v)) for v in region_start)
    stop = img if region_stop is None else tuple(min(img[a], int(region_stop[a])) for a in range(3))
    if any(stop[a] <= start[a] for a in range(3)):
        raise ValueError(f"Empty lazy inference region: start={start}, stop={stop}")
    osz = tuple(stop[a] - start[a] for a in range(3))
    ov = tuple(float(v) for v in overlap) if isinstance(overlap, (list, tuple)) else (float(overlap),) * 3
    recs = lazy_window_records(img, roi, ov, start, stop, snap_to_edge)[rank::world_size]
    if not recs:
        raise RuntimeError("No lazy sliding-window patches were generated" + (f" on rank {rank}" if world_size > 1 else "."))
    wmap = W.build_sliding_importance_map(roi, mode=mode, device=dev, dtype=output_dtype)
    wmap = W.apply_border_mask(wmap, list(border_mask))
    value = None
    weight = torch.zeros((1, 1, *osz), device=dev, dtype=output_dtype)
    for b0 in range(0, len(recs), max(1, int(sw_batch_size))):
        chunk = recs[b0:b0 + sw_batch_size]
        batch = W._extract_starts(vol, [r[0] for r in chunk], roi, padding_mode, cval)
        with torch.no_grad():
            pred = network(batch.float())
        pred = pred.detach().to(device=dev, dtype=output_dtype).contiguous()
        if value is None:
            value = torch.zeros((1, int(pred.shape[1]), *osz), device=dev, dtype=output_dtype)
        for i, (_, plo, box, olo) in enumerate(chunk):
            W._accumulate_window(pred[i], wmap, value, weight, roi, osz, plo, olo, box)
    if accumulator_reduce is not None:
        reduced = accumulator_reduce(value, weight)
        if reduced is None:
            return torch.empty(0, device=dev)
        value, weight = reduced
    if not normalize:
        return value, weight
    return W.normalize_weighted_accumulator(value, weight)


def lazy_predict_region(volume, network, *, region_start, region_stop, **kw):
    """``lazy.py:1261-1293`` on an in-memory volume."""
    return lazy_sliding_window(volume, network, region_start=region_start, region_stop=region_stop, **kw)


def lazy_predict_volume(volume, network, **kw):
    """``lazy.py:1295-1334`` on an in-memory volume."""
    return lazy_sliding_window(volume, network, region_start=None, region_stop=None, **kw)


__all__ = ["lazy_window_records", "lazy_sliding_window", "lazy_predict_region", "lazy_predict_volume"]
