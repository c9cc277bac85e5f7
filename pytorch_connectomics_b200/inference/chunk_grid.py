"""Config-side helpers of chunked inference under the reference's names (``connectomics/inference/chunk_grid.py``):
crop pads, the chunk shape of ``inference.chunking``, output mode / backend validation, the h5 chunking of the streamed
output, and the global prediction crop that DeepEM-style affinities need.  Pure host integer logic; the chunk grid itself
(``ChunkRef`` / ``build_chunk_grid`` / ``resolve_halo_region``) lives in :mod:`.chunked`.
"""

from __future__ import annotations

from typing import Any, List, Optional, Sequence, Tuple

from .tta import resolve_channel_indices
from .tta_affinity import normalize_affinity_mode, resolve_affinity_channel_groups_from_cfg, resolve_affinity_mode_from_cfg

Pad3 = Tuple[Tuple[int, int], Tuple[int, int], Tuple[int, int]]
_NO_PAD: Pad3 = ((0, 0), (0, 0), (0, 0))


def _node(obj: Any, *path: str, default: Any = None) -> Any:
    for key in path:
        if obj is None:
            return default
        obj = obj.get(key, None) if isinstance(obj, dict) else getattr(obj, key, None)
    return default if obj is None else obj


def normalize_crop_pad(value: Any) -> Pad3:
    """``chunk_grid.py:22-30`` — ``None`` / 3 symmetric / 6 (before, after) values -> per-axis (before, after)."""
    if value is None or (hasattr(value, "__len__") and len(value) == 0):
        return _NO_PAD
    vals = [int(v) for v in value]
    if len(vals) == 3:
        return tuple((v, v) for v in vals)  # type: ignore[return-value]
    if len(vals) == 6:
        return tuple((vals[2 * a], vals[2 * a + 1]) for a in range(3))  # type: ignore[return-value]
    raise ValueError(f"inference.model.crop_pad must have length 3 or 6, got {value!r}")


def compute_affinity_crop_pad(offsets: Sequence[Sequence[int]], *, affinity_mode: str = "deepem") -> Tuple[Tuple[int, int], ...]:
    """``data/processing/affinity.py:291-315`` — the border where an affinity with these offsets has no valid partner voxel:
    DeepEM-style edges look backwards (positive offset -> leading border), the other convention forwards."""
    if not offsets:
        return tuple()
    deepem = normalize_affinity_mode(affinity_mode) == "deepem"
    ndim = len(offsets[0])
    if any(len(o) != ndim for o in offsets):
        raise ValueError(f"Mixed affinity offset dimensions are not supported: {offsets!r}")
    lead = [max([0] + [int(o[a]) if deepem else -int(o[a]) for o in offsets]) for a in range(ndim)]
    trail = [max([0] + [-int(o[a]) if deepem else int(o[a]) for o in offsets]) for a in range(ndim)]
    return tuple((lead[a], trail[a]) for a in range(ndim))


def resolve_selected_affinity_offsets(cfg: Any) -> List[Tuple[int, int, int]]:
    """``chunk_grid.py:33-54`` — offsets of the affinity channels that survive ``inference.model.select_channel``."""
    groups = resolve_affinity_channel_groups_from_cfg(cfg)
    if not groups:
        return []
    per_channel: List[Optional[Tuple[int, int, int]]] = [None] * max(hi for (_, hi), _ in groups)
    for (lo, hi), offs in groups:
        for ch, off in zip(range(lo, hi), offs):
            per_channel[ch] = tuple(int(v) for v in off)  # type: ignore[assignment]
    select = _node(cfg, "inference", "model", "select_channel")
    if select is not None:
        keep = resolve_channel_indices(select, num_channels=len(per_channel), context="inference.model.select_channel")
        per_channel = [per_channel[i] for i in keep]
    return [o for o in per_channel if o is not None]


def resolve_global_prediction_crop(cfg: Any) -> Pad3:
    """``chunk_grid.py:57-77`` — the user's ``inference.model.crop_pad`` plus, for DeepEM affinities, the invalid border."""
    user = normalize_crop_pad(_node(cfg, "inference", "model", "crop_pad"))
    aff: Sequence[Tuple[int, int]] = _NO_PAD
    if resolve_affinity_mode_from_cfg(cfg) == "deepem":
        offs = resolve_selected_affinity_offsets(cfg)
        if offs:
            aff = compute_affinity_crop_pad(offs, affinity_mode="deepem")
    return tuple((int(user[a][0]) + int(aff[a][0]), int(user[a][1]) + int(aff[a][1])) for a in range(3))  # type: ignore[return-value]


def validate_chunked_output_format(cfg: Any) -> None:
    """``chunk_grid.py:80-87``"""
    backend = str(_node(cfg, "inference", "save_backend", default="h5")).lower()
    if backend not in ("h5", "hdf5"):
        raise ValueError("Chunked inference writes a single streamed HDF5 output only; "
                         f"unsupported inference.save_backend={backend!r}.")


def resolve_chunk_shape(cfg: Any, final_shape: Sequence[int]) -> Tuple[int, int, int]:
    """``chunk_grid.py:90-98`` — ``inference.chunking.chunk_size`` clipped to the volume; ``axes='z'`` = full xy slabs."""
    chunking = cfg.inference.chunking
    size = tuple(int(v) for v in chunking.chunk_size)
    axes = str(getattr(chunking, "axes", "all")).lower()
    if axes == "z":
        return (size[0], int(final_shape[1]), int(final_shape[2]))
    if axes != "all":
        raise ValueError("inference.chunking.axes must be 'all' or 'z'")
    return tuple(min(size[a], int(final_shape[a])) for a in range(3))  # type: ignore[return-value]


def resolve_h5_spatial_chunks(spatial_shape: Sequence[int]) -> Tuple[int, int, int]:
    """``chunk_grid.py:101-103`` — 64^3 h5 chunks, smaller where the volume is"""
    return tuple(min(int(spatial_shape[a]), 64) for a in range(3))  # type: ignore[return-value]


def resolve_chunk_output_mode(cfg: Any) -> str:
    """``chunk_grid.py:106-111``"""
    mode = str(getattr(cfg.inference.chunking, "output_mode", "decoded")).lower()
    if mode not in ("decoded", "raw_prediction"):
        raise ValueError("inference.chunking.output_mode must be 'decoded' or 'raw_prediction'.")
    return mode


__all__ = ["normalize_crop_pad", "resolve_selected_affinity_offsets", "resolve_global_prediction_crop",
           "validate_chunked_output_format", "resolve_chunk_shape", "resolve_h5_spatial_chunks", "resolve_chunk_output_mode",
           "compute_affinity_crop_pad"]
