"""Lazy-grid sliding-window inference over a bounded region — ``connectomics/inference/lazy.py`` on the B200 engine.

Two layers:

* the reference's own seam, callable unchanged by its callers (``inference/chunked.py:541-551``,
  ``tests/unit/test_lazy_inference.py``): ``lazy_predict_region(cfg, forward_fn, image_path, *, region_start,
  region_stop, mask_path, mask_align_to_image, device, requested_head)`` and ``lazy_predict_volume(cfg, forward_fn,
  image_path, *, mask_path, ...)`` (``lazy.py:1261-1334``), driven by ``cfg.inference.sliding_window`` exactly like
  ``_lazy_sliding_window`` (``:986-1258``): roi / overlap / sw_batch_size / blending / padding_mode / cval /
  snap_to_edge / ``target_context`` (``:368-419``) / border_mask / ``inference.model.output_dtype``.
  ``image_path`` is anything :func:`build_accessor` understands: a path to a ``.npy`` volume (memory-mapped; h5/zarr only
  if those packages import), a ``torch.Tensor`` / ``numpy`` array, or any object with the accessor protocol of
  ``LazyVolumeAccessor`` (``padded_spatial_shape``, ``channel_count``, ``read_patch(location, patch_size, *,
  outer_pad_mode, outer_pad_value)``, ``lazy.py:456-918``).  The reference's test-time transforms (resize, transpose,
  normalisation, tile mosaics) are data-pipeline work outside this path: accessors here serve the stored
  voxels.

* ``lazy_sliding_window(volume, network, ...)`` — the tile loop itself on a tensor: window offsets with face-centred
  boundary windows (``lazy.py:269-334``, negative starts) from ``pcb_sw_plan(PCB_GRID_LAZY | PCB_GRID_LAZY_SNAP)``,
  filtered to the windows that intersect the region (``:337-365``); patches cast to fp32, run through ``network``, cast
  to the output dtype (``:1184-1206``); only the part of each window inside the region is accumulated
  (``:1077-1102,1216-1227``); ``rank``/``world_size`` shard the records as ``records[rank::world_size]`` (``:1104``)
  and ``accumulator_reduce`` sees the un-normalised accumulators (``:1241-1249``).
"""

from __future__ import annotations

import os
from typing import Any, Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib as L
from . import window as W


# ----------------------------------------------------------------------------- accessors (lazy.py:456-918 protocol)
def _pad_channel_first(array: np.ndarray, pads, *, mode: str, constant_value: float = 0.0) -> np.ndarray:
    """``lazy.py:237-256``: numpy padding of a [C, z, y, x] crop (``replicate`` -> ``edge``; reflect on a size-1 axis ->
    edge)."""
    if not any(b > 0 or a > 0 for b, a in pads):
        return array
    np_mode = "edge" if str(mode).lower() == "replicate" else str(mode).lower()
    width = [(0, 0)] + [(int(b), int(a)) for b, a in pads]
    if np_mode == "constant":
        return np.pad(array, width, mode="constant", constant_values=constant_value)
    if np_mode == "circular":
        np_mode = "wrap"
    if np_mode == "reflect" and any(s <= 1 for s in array.shape[1:]):
        np_mode = "edge"
    return np.pad(array, width, mode=np_mode)


class ArrayVolumeAccessor:
    """Accessor over an in-memory or memory-mapped ``[C, D, H, W]`` / ``[D, H, W]`` array (numpy or torch).
    Same surface the tile loop uses on ``LazyVolumeAccessor``: context manager, ``padded_spatial_shape``,
    ``channel_count``, ``read_patch`` (fp32 ``[C, *patch_size]``, outer padding as ``lazy.py:852-904``), ``load_full``."""

    def __init__(self, data, *, kind: str = "image", binarize: bool = False, threshold: float = 0.0,
                 context_pad: Sequence[Sequence[int]] = ((0, 0), (0, 0), (0, 0)), context_pad_mode: str = "constant",
                 transpose_axes: Sequence[int] = (), scale_factors: Optional[Sequence[float]] = None,
                 layout: str = "channel_first", normalize_mode: str = "none", clip_percentile_low: float = 0.0,
                 clip_percentile_high: float = 1.0):
        # data.image_transform.normalize (lazy.py:895-902): intensity normalisation of every PATCH an image accessor hands out
        self.normalize_mode = str(normalize_mode or "none")
        self.clip_percentile_low, self.clip_percentile_high = float(clip_percentile_low), float(clip_percentile_high)
        if isinstance(data, torch.Tensor):
            if data.dim() == 5:
                if data.shape[0] != 1:
                    raise ValueError(f"expected a [1, C, D, H, W] volume, got {tuple(data.shape)}")
                data = data[0]
            if data.dim() == 3:
                data = data.unsqueeze(0)
            self._tensor: Optional[torch.Tensor] = data
            self._array = None
        else:
            arr = data
            if arr.ndim == 5 and arr.shape[0] == 1:
                arr = arr[0]
            if arr.ndim == 4 and layout == "infer":
                # lazy.py:572-586: a 4-D DATASET says nothing about where its channel axis is — the reference takes the
                # smallest axis (first / last / second), spatial axes keep their order
                smallest = int(np.argmin(arr.shape))
                if smallest == 3:
                    arr = np.moveaxis(arr, -1, 0)
                elif smallest == 1:
                    arr = np.transpose(arr, (1, 0, 2, 3))
            if arr.ndim == 3:
                arr = arr[None]
            self._tensor = None
            self._array = arr
        ref = self._tensor if self._tensor is not None else self._array
        if len(ref.shape) != 4:
            raise ValueError(f"volume must be [C, D, H, W] or [D, H, W]; got shape {tuple(ref.shape)}")
        self.kind = kind
        self.binarize = bool(binarize)
        self.threshold = float(threshold)
        self.channel_count = int(ref.shape[0])
        # data_transform.val_transpose (lazy.py:922): the logical volume is the stored one with its spatial axes permuted
        self.transpose_axes = tuple(int(a) for a in (transpose_axes or ()))
        if self.transpose_axes and sorted(self.transpose_axes) != [0, 1, 2]:
            raise ValueError(f"transpose_axes must be a permutation of (0, 1, 2), got {self.transpose_axes}")
        stored = tuple(int(v) for v in ref.shape[1:])
        self.logical_spatial_shape = tuple(stored[a] for a in self.transpose_axes) if self.transpose_axes else stored
        # test-time resize (lazy.py:422-453,508-515): the transformed volume is the logical one resampled by these factors —
        # trilinear with align_corners for images, nearest for masks / labels — evaluated per read, never materialised
        self.scale_factors = tuple(float(v) for v in scale_factors) if scale_factors else None
        if self.scale_factors is not None and any(f <= 0 for f in self.scale_factors):
            raise ValueError(f"scale factor must be positive, got {min(self.scale_factors)}.")
        self.transformed_spatial_shape = self.logical_spatial_shape if self.scale_factors is None else tuple(
            max(1, int(np.floor(float(n) * f + 1e-6))) for n, f in zip(self.logical_spatial_shape, self.scale_factors))
        # data_transform.pad_size / pad_mode (lazy.py:924-929): the context border the test-time transform adds around
        # the volume; window coordinates, the reference shape and the prediction all live in the PADDED frame
        self.context_pad = tuple((int(b), int(a)) for b, a in context_pad)
        mode = str(context_pad_mode).lower()
        self.context_pad_mode = "edge" if mode == "replicate" else mode
        if self.context_pad_mode not in ("constant", "reflect", "edge"):
            raise ValueError(f"Unsupported context pad mode '{context_pad_mode}'.")
        self.padded_spatial_shape = tuple(self.transformed_spatial_shape[a] + self.context_pad[a][0] + self.context_pad[a][1]
                                          for a in range(3))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self) -> None:
        pass

    def as_tensor(self) -> Optional[torch.Tensor]:
        """``[1, C, D, H, W]`` view when the whole volume is a tensor (device-resident fast path), else ``None``."""
        plain = not self.binarize and not self.transpose_axes and not any(b or a for b, a in self.context_pad) \
            and self.scale_factors is None and not (self.kind == "image" and self.normalize_mode != "none")
        if self._tensor is not None and plain:
            return self._tensor.unsqueeze(0)
        return None

    def _crop(self, lo, hi) -> np.ndarray:
        """[C, *box] of the TRANSFORMED (transposed, resized, unpadded) volume (``lazy.py:738-792``)"""
        if self.scale_factors is None:
            return self._raw_crop(lo, hi)
        shape = tuple(int(hi[a]) - int(lo[a]) for a in range(3))
        if any(v <= 0 for v in shape):
            return np.zeros((self.channel_count, *shape), dtype=np.float32)
        nearest = self.kind in ("label", "mask")
        coords = []
        for a in range(3):
            idx = np.arange(int(lo[a]), int(hi[a]), dtype=np.float32)
            n_in, n_out = int(self.logical_spatial_shape[a]), int(self.transformed_spatial_shape[a])
            if n_in <= 1 or n_out <= 1:
                c = np.zeros_like(idx)
            elif nearest:
                c = np.clip(np.floor(idx * float(n_in) / float(n_out)), 0, n_in - 1).astype(np.float32)
            else:                                   # bilinear, align_corners=True
                c = (idx * float(n_in - 1) / float(n_out - 1)).astype(np.float32)
            coords.append(c)
        r0 = tuple(int(np.floor(float(c.min()))) for c in coords)
        r1 = tuple(min(int(self.logical_spatial_shape[a]), int(np.ceil(float(coords[a].max()))) + 1) for a in range(3))
        raw = self._raw_crop(r0, r1)
        local = [coords[a] - float(r0[a]) for a in range(3)]
        if nearest:
            out = raw
            for a in range(3):
                out = np.take(out, local[a].astype(np.int64), axis=a + 1)
            return out.astype(np.float32, copy=False)
        grid_axes = [(np.zeros_like(local[a]) if raw.shape[a + 1] <= 1 else (2.0 * local[a]) / float(raw.shape[a + 1] - 1) - 1.0)
                     .astype(np.float32) for a in range(3)]
        zz, yy, xx = np.meshgrid(grid_axes[0], grid_axes[1], grid_axes[2], indexing="ij")
        grid = torch.from_numpy(np.stack([xx, yy, zz], axis=-1)).unsqueeze(0)
        sampled = torch.nn.functional.grid_sample(torch.from_numpy(np.ascontiguousarray(raw)).unsqueeze(0), grid, mode="bilinear",
                                                  padding_mode="zeros", align_corners=True)
        return sampled.squeeze(0).numpy().astype(np.float32, copy=False)

    def _raw_crop(self, lo, hi) -> np.ndarray:
        """[C, *box] of the LOGICAL (transposed, not resized) volume"""
        if self.transpose_axes:            # logical axis i is stored axis transpose_axes[i]
            raw = [None, None, None]
            for i, a in enumerate(self.transpose_axes):
                raw[a] = slice(int(lo[i]), int(hi[i]))
            sl = (slice(None), *raw)
        else:
            sl = (slice(None),) + tuple(slice(int(a), int(b)) for a, b in zip(lo, hi))
        if self._tensor is not None:
            out = self._tensor[sl].detach().to("cpu", torch.float32).numpy()
        else:
            out = np.asarray(self._array[sl], dtype=np.float32)
        return np.transpose(out, (0, *[a + 1 for a in self.transpose_axes])) if self.transpose_axes else out

    def _read_padded_inner(self, lo, hi) -> np.ndarray:
        """``lazy.py:794-850``: the box [lo, hi) of the PADDED volume — every padded index is mapped to the source index the
        context padding takes its value from (reflect without edge repeat, edge clamp, or zero outside for constant)."""
        mapped, valid, b0, b1 = [], [], [], []
        for a in range(3):
            idx = np.arange(int(lo[a]), int(hi[a]), dtype=np.int64) - self.context_pad[a][0]
            n = self.transformed_spatial_shape[a]
            if self.context_pad_mode == "reflect" and n > 1:
                period = 2 * n - 2
                m = np.abs(idx) % period
                m = np.where(m < n, m, period - m)
                ok = np.ones(idx.shape, dtype=bool)
            elif self.context_pad_mode == "reflect":
                m, ok = np.zeros_like(idx), np.ones(idx.shape, dtype=bool)
            else:
                m = np.clip(idx, 0, max(n - 1, 0))
                ok = np.ones(idx.shape, dtype=bool) if self.context_pad_mode == "edge" else (idx >= 0) & (idx < n)
            mapped.append(m); valid.append(ok)
            b0.append(int(m.min()) if m.size else 0); b1.append(int(m.max()) + 1 if m.size else 0)
        region = self._crop(b0, b1)
        out = region
        for a in range(3):
            out = np.take(out, mapped[a] - b0[a], axis=a + 1)
        if self.context_pad_mode == "constant":
            out = out * (valid[0][:, None, None] & valid[1][None, :, None] & valid[2][None, None, :])[None].astype(out.dtype)
        return out

    def read_patch(self, location, patch_size, *, outer_pad_mode: str, outer_pad_value: float) -> np.ndarray:
        start = tuple(int(v) for v in location)
        size = tuple(int(v) for v in patch_size)
        end = tuple(start[i] + size[i] for i in range(3))
        lo = tuple(max(0, start[i]) for i in range(3))
        hi = tuple(min(self.padded_spatial_shape[i], end[i]) for i in range(3))
        if any(hi[i] <= lo[i] for i in range(3)):          # a box entirely outside: the reference pads an EMPTY crop (lazy.py:866-867),
            inner = np.zeros((self.channel_count, 0, 0, 0), dtype=np.float32)     # which numpy only accepts for constant padding
        else:
            inner = self._read_padded_inner(lo, hi) if (any(b or a for b, a in self.context_pad)) else self._crop(lo, hi)
        pads = [(max(0, -start[i]), max(0, end[i] - self.padded_spatial_shape[i])) for i in range(3)]
        patch = _pad_channel_first(inner, pads, mode=outer_pad_mode, constant_value=outer_pad_value)
        if self.binarize:
            patch = (patch > self.threshold).astype(np.float32, copy=False)
        if self.kind == "image" and self.normalize_mode != "none":
            patch = normalize_patch(patch, self.normalize_mode, self.clip_percentile_low, self.clip_percentile_high)
        return patch.astype(np.float32, copy=False)

    def load_full(self) -> np.ndarray:
        """``lazy.py:906-918``: the transformed volume WITHOUT the context border"""
        full = self._crop((0, 0, 0), self.transformed_spatial_shape)
        return (full > self.threshold).astype(np.float32) if self.binarize else full


def normalize_patch(patch: np.ndarray, mode: str, clip_low: float = 0.0, clip_high: float = 1.0) -> np.ndarray:
    """``smart_normalize`` (``data/augmentation/augment_ops.py:552-610``) as the lazy accessor applies it (``lazy.py:895-902``):
    statistics of THIS patch (outer padding included), never of the volume.  Percentile clip first (fractions in [0, 1]), then
    ``"normal"`` (z-score unless std <= 1e-8), ``"0-1"`` (min-max unless flat) or ``"divide-K"``.  numpy's own reductions in
    the reference's order, so the float results are the reference's bit for bit."""
    divisor = None
    if mode.startswith("divide-"):
        try:
            divisor = float(mode.split("-", 1)[1])
        except ValueError as exc:
            raise ValueError(f"Invalid divide mode '{mode}'. Format should be 'divide-K' where K is a number "
                             "(e.g., 'divide-255').") from exc
        mode = "divide"
    out = np.array(patch, copy=True)
    if clip_low > 0.0 or clip_high < 1.0:
        bounds = (np.percentile(out, clip_low * 100), np.percentile(out, clip_high * 100))
        out = np.clip(out, *bounds)
    if mode == "normal":
        mu, sigma = out.mean(), out.std()
        return (out - mu) / sigma if sigma > 1e-8 else out
    if mode == "0-1":
        lo, hi = out.min(), out.max()
        return (out - lo) / (hi - lo) if hi > lo else out
    if mode == "divide":
        if divisor is None or divisor == 0.0:       # the accessor never passes a divide_value: plain "divide" is an error there too
            raise ValueError("smart_normalize mode='divide' requires a non-zero divide_value (or use 'divide-K' form to "
                             "embed the divisor in the mode string).")
        return out / divisor
    if mode == "none":
        return out
    raise ValueError(f"Unknown smart_normalize mode '{mode}'. Expected 'none', 'normal', '0-1', 'divide', or 'divide-K'.")


def _get_padsize(pad_size, ndim: int = 3):
    """``data/processing/misc.py:20-41`` get_padsize: int | [p] | [pz, py, px] | [z0, z1, y0, y1, x0, x1] -> per-axis pairs"""
    if isinstance(pad_size, int):
        return tuple((pad_size, pad_size) for _ in range(ndim))
    vals = list(pad_size)
    if len(vals) not in (1, ndim, 2 * ndim):
        raise ValueError(f"pad_size length must be 1, {ndim}, or {2 * ndim}, got {len(vals)}")
    if len(vals) == 1:
        return tuple((vals[0], vals[0]) for _ in range(ndim))
    if len(vals) == ndim:
        return tuple((v, v) for v in vals)
    return tuple((vals[2 * i], vals[2 * i + 1]) for i in range(ndim))


_ACCESSOR_FACTORIES: List[Callable[..., Any]] = []


def register_accessor_factory(factory: Callable[..., Any]) -> None:
    """``factory(cfg, source, kind=..., mode=...) -> accessor | None``; consulted before the built-in sources (the place a
    deployment plugs the reference's h5/zarr/tiff ``LazyVolumeAccessor`` in)."""
    _ACCESSOR_FACTORIES.insert(0, factory)


def build_accessor(cfg, source, *, kind: str = "image", mode: str = "test"):
    """``lazy.py:920-959`` seam: image/mask source -> accessor."""
    for factory in _ACCESSOR_FACTORIES:
        acc = factory(cfg, source, kind=kind, mode=mode)
        if acc is not None:
            return acc
    binarize, threshold = False, 0.0
    data_cfg = getattr(cfg, "data", None)
    dt = getattr(data_cfg, "data_transform", None)
    if kind == "mask":
        mask_cfg = getattr(data_cfg, "mask_transform", None) or dt
        binarize = bool(getattr(mask_cfg, "binarize", False))
        threshold = float(getattr(mask_cfg, "threshold", 0.0))
    if hasattr(source, "read_patch") and hasattr(source, "padded_spatial_shape"):
        return source
    # lazy.py:922-929: transpose, context border (image: data_transform.pad_mode, default reflect; mask: zeros)
    kw = dict(kind=kind, binarize=binarize, threshold=threshold,
              transpose_axes=tuple(getattr(dt, "val_transpose", None) or ()),
              context_pad=_get_padsize(getattr(dt, "pad_size", [0, 0, 0])) if kind in ("image", "mask") else ((0, 0),) * 3,
              context_pad_mode=getattr(dt, "pad_mode", "reflect") if kind == "image" else "constant")
    # lazy.py:422-453: test-time resize factors (data_transform.resize is a SIZE relative to dataloader.patch_size)
    factors = None
    target = getattr(dt, "resize", None) if mode in ("test", "tune") else None
    if target:                                   # data_transform.resize is a SIZE: factor = size / dataloader.patch_size
        patch = getattr(getattr(data_cfg, "dataloader", None), "patch_size", None)
        if not (patch and len(patch) == len(target) and all(float(v) > 0 for v in patch)):
            raise ValueError("Lazy sliding-window inference requires data.dataloader.patch_size when "
                             "data_transform.resize is configured.")
        factors = [float(o) / float(i) for o, i in zip(target, patch)]
    elif kind in ("image", "label"):
        factors = getattr(getattr(data_cfg, "image_transform", None), "resize", None)
    elif kind == "mask":
        factors = getattr(getattr(data_cfg, "mask_transform", None) or dt, "resize", None)
    kw["scale_factors"] = tuple(float(v) for v in factors) if factors else None
    if kind == "image":                          # lazy.py:931-937
        it = getattr(data_cfg, "image_transform", None)
        kw["normalize_mode"] = getattr(it, "normalize", "none") or "none"
        kw["clip_percentile_low"] = float(getattr(it, "clip_percentile_low", 0.0))
        kw["clip_percentile_high"] = float(getattr(it, "clip_percentile_high", 1.0))
    if isinstance(source, (torch.Tensor, np.ndarray)):
        return ArrayVolumeAccessor(source, **kw)
    path = os.fspath(source)
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npy":
        return ArrayVolumeAccessor(np.load(path, mmap_mode="r"), layout="infer", **kw)
    if ext in (".h5", ".hdf5"):
        try:
            import h5py  # noqa: F401
        except ImportError as exc:
            raise RuntimeError(f"pcb200 lazy inference: reading {path} needs h5py, which is not installed; pass a .npy "
                               "volume, an array, or register_accessor_factory(...)") from exc
        f = h5py.File(path, "r")
        ds = f[next(iter(f.keys()))]
        if ds.ndim == 4 and int(np.argmin(ds.shape)) in (1, 3):      # channel-second / channel-last: re-laid out in memory;
            ds = ds[...]                                             # channel-first and 3-D datasets stay lazy (sliced per read)
        return ArrayVolumeAccessor(ds, layout="infer", **kw)
    raise ValueError(f"pcb200 lazy inference: unsupported volume source {source!r}; expected a .npy path, a tensor/array "
                     "or an accessor object (register_accessor_factory adds formats).")


def get_lazy_image_reference_shape(cfg, image_path, *, mode: str = "test") -> Tuple[int, ...]:
    """``lazy.py:962-978``."""
    with build_accessor(cfg, image_path, kind="image", mode=mode) as acc:
        patch = getattr(getattr(getattr(cfg, "data", None), "dataloader", None), "patch_size", None)
        shape = getattr(acc, "transformed_spatial_shape", acc.padded_spatial_shape)
        if patch and any(int(shape[i]) < int(patch[i]) for i in range(3)):
            raise ValueError("Lazy sliding-window inference currently requires the transformed test volume "
                             "to be at least as large as data.dataloader.patch_size in every axis. "
                             f"Got transformed_shape={tuple(int(v) for v in shape)}, "
                             f"patch_size={tuple(int(v) for v in patch)}.")
        return (1, int(acc.channel_count), *tuple(int(v) for v in acc.padded_spatial_shape))


# ----------------------------------------------------------------------------- target context (lazy.py:368-419)
def _resolve_target_context(sliding_cfg, roi_size: Sequence[int]) -> Tuple[int, int, int]:
    context_cfg = list(getattr(sliding_cfg, "target_context", []) or [])
    if not context_cfg:
        return (0, 0, 0)
    if len(context_cfg) == 1:
        context_cfg = context_cfg * 3
    if len(context_cfg) != 3:
        raise ValueError("inference.sliding_window.target_context must have length 1 or 3, "
                         f"got {context_cfg}.")
    context = tuple(int(v) for v in context_cfg)
    if any(v < 0 for v in context):
        raise ValueError(f"inference.sliding_window.target_context values must be non-negative, got {context}.")
    if len(tuple(roi_size)) != 3:
        raise ValueError("Lazy sliding-window target_context currently supports 3D only.")
    return context


def _crop_prediction_to_roi(prediction: torch.Tensor, *, roi_size: Sequence[int], target_context: Sequence[int],
                            scope: str) -> torch.Tensor:
    spatial = tuple(int(v) for v in prediction.shape[2:])
    roi = tuple(int(v) for v in roi_size)
    context = tuple(int(v) for v in target_context)
    if not any(context):
        if spatial != roi:
            raise RuntimeError(f"{scope} requires model predictions to have the same spatial shape as the "
                               f"sliding-window ROI. Got prediction.shape={tuple(prediction.shape)} and "
                               f"roi_size={roi}.")
        return prediction
    expected = tuple(roi[a] + 2 * context[a] for a in range(3))
    if spatial != expected:
        raise RuntimeError(f"{scope} with target_context={context} expected prediction spatial shape "
                           f"{expected}, got {spatial}.")
    sl = [slice(None), slice(None)] + [slice(context[a], context[a] + roi[a]) for a in range(3)]
    return prediction[tuple(sl)]


# ----------------------------------------------------------------------------- grid records
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


# ----------------------------------------------------------------------------- the tile loop
def _lazy_tile_loop(read_batch: Callable[[list], torch.Tensor], predict: Callable[[torch.Tensor, list], torch.Tensor],
                    *, image_size, roi, overlap, mode, region_start, region_stop, snap_to_edge, sw_batch_size,
                    output_dtype, border_mask, rank, world_size, accumulator_reduce, dev, normalize, target_context,
                    what: str = "", shard_validator=None):
    img = tuple(int(v) for v in image_size)
    if any(img[a] < roi[a] for a in range(3)):
        raise ValueError("Lazy sliding-window inference requires the transformed test volume to be at least as large "
                         f"as the ROI in every axis. Got bounds_shape={img}, roi_size={roi}.")
    start = (0, 0, 0) if region_start is None else tuple(max(0, int(v)) for v in region_start)
    stop = img if region_stop is None else tuple(min(img[a], int(region_stop[a])) for a in range(3))
    if any(stop[a] <= start[a] for a in range(3)):
        raise ValueError(f"Empty lazy inference region: start={start}, stop={stop}")
    osz = tuple(stop[a] - start[a] for a in range(3))
    ov = tuple(float(v) for v in overlap) if isinstance(overlap, (list, tuple)) else (float(overlap),) * 3
    every = lazy_window_records(img, roi, ov, start, stop, snap_to_edge)
    recs = every[rank::world_size]
    if shard_validator is not None and world_size > 1:     # collective: every rank raises when any shard is empty
        shard_validator(len(recs), len(every))
    if not recs:
        raise RuntimeError(f"No lazy sliding-window patches were generated{what}" +
                           (f" on rank {rank}" if world_size > 1 else "."))
    wmap = W.build_sliding_importance_map(roi, mode=mode, device=dev, dtype=output_dtype)
    wmap = W.apply_border_mask(wmap, list(border_mask))
    value = None
    weight = torch.zeros((1, 1, *osz), device=dev, dtype=output_dtype)
    bs = max(1, int(sw_batch_size))
    for b0 in range(0, len(recs), bs):
        chunk = recs[b0:b0 + bs]
        batch = read_batch([r[0] for r in chunk])
        with torch.no_grad():
            pred = predict(batch, chunk)
        if not isinstance(pred, torch.Tensor):
            raise ValueError(f"lazy sliding-window: `network` must return a torch.Tensor; got {type(pred).__name__}.")
        pred = _crop_prediction_to_roi(pred, roi_size=roi, target_context=target_context,
                                       scope="Lazy sliding-window inference")
        W.check_network_output(pred, len(chunk), None if value is None else int(value.shape[1]), roi,
                               "lazy sliding-window")
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


def lazy_sliding_window(volume: torch.Tensor, network: Callable[[torch.Tensor], torch.Tensor], *, roi_size,
                        overlap=0.5, mode: str = "bump", padding_mode: str = "constant", cval: float = 0.0,
                        region_start: Optional[Sequence[int]] = None, region_stop: Optional[Sequence[int]] = None,
                        snap_to_edge: bool = False, sw_batch_size: int = 1, output_dtype: torch.dtype = torch.float32,
                        border_mask: Sequence[int] = (), rank: int = 0, world_size: int = 1,
                        accumulator_reduce=None, device=None, normalize: bool = True,
                        target_context: Sequence[int] = (0, 0, 0)):
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
    ctx = tuple(int(v) for v in target_context)
    read_size = tuple(roi[a] + 2 * ctx[a] for a in range(3))

    def read_batch(starts):
        shifted = [tuple(s[a] - ctx[a] for a in range(3)) for s in starts]
        return W._extract_starts(vol, shifted, read_size, padding_mode, cval)

    return _lazy_tile_loop(read_batch, lambda b, _c: network(b.float()), image_size=vol.shape[-3:], roi=roi,
                           overlap=overlap, mode=mode, region_start=region_start, region_stop=region_stop,
                           snap_to_edge=snap_to_edge, sw_batch_size=sw_batch_size, output_dtype=output_dtype,
                           border_mask=border_mask, rank=rank, world_size=world_size,
                           accumulator_reduce=accumulator_reduce, dev=dev, normalize=normalize, target_context=ctx)


# ----------------------------------------------------------------------------- the reference's seam
from .model_outputs import pick_inference_output  # noqa: E402


def _dist_context():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return True, dist.get_rank(), dist.get_world_size()
    return False, 0, 1


def _lazy_sliding_window_cfg(cfg, forward_fn, image_path, *, region_start, region_stop, mask_path, mask_align_to_image,
                             device, requested_head, enable_distributed_window_sharding: bool, accumulator_reduce=None,
                             shard_validator=None):
    """``lazy.py:986-1258`` with the B200 kernels for map / accumulate / normalise.  Accumulators live on ``device``
    (the reference keeps them on the CPU); the result is returned on the CPU as the reference does."""
    # lazy.py:1038: the reference runs every patch batch through TTAPredictor(cfg, None, forward_fn).predict — views,
    # activations, channel selection and the mask per PATCH.  The predictor is only built when the config asks for any of
    # that; a plain forward keeps the direct route (head selection + mask product).
    tta_cfg = getattr(getattr(cfg, "inference", None), "test_time_augmentation", None)
    inf_model = getattr(getattr(cfg, "inference", None), "model", None)
    predictor = None
    if (tta_cfg is not None and getattr(tta_cfg, "enabled", False)) or getattr(inf_model, "channel_activations", None) \
            or getattr(inf_model, "select_channel", None) is not None:
        from .tta_predictor import TTAPredictor
        predictor = TTAPredictor(cfg, None, forward_fn)
    roi_size = W.resolve_inferer_roi_size(cfg)
    if roi_size is None:
        raise ValueError("Lazy sliding-window inference requires inference.sliding_window.window_size "
                         "or model.output_size to be configured.")
    if len(roi_size) != 3:
        raise ValueError(f"Lazy sliding-window inference currently supports 3D only, got {roi_size}.")
    roi = tuple(int(v) for v in roi_size)
    overlap = W.resolve_inferer_overlap(cfg, roi)
    sc = getattr(getattr(cfg, "inference", None), "sliding_window", None)
    loader = getattr(getattr(cfg, "data", None), "dataloader", None)
    sw_batch_size = max(1, int(getattr(sc, "sw_batch_size", None) or getattr(loader, "batch_size", 1) or 1))
    mode = str(getattr(sc, "blending", "bump")).strip().lower()
    pad_mode = getattr(sc, "padding_mode", "constant")
    cval = float(getattr(sc, "cval", 0.0))
    snap = bool(getattr(sc, "snap_to_edge", False))
    ctx = _resolve_target_context(sc, roi)
    read_size = tuple(roi[a] + 2 * ctx[a] for a in range(3))
    dev = torch.device(device)
    if dev.type == "cuda" and dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    W._device_or_raise(dev)
    _, rank, world = _dist_context()
    if not enable_distributed_window_sharding:
        rank, world = 0, 1
    border_mask = W.resolve_border_mask(cfg, 3)
    output_dtype = W.resolve_model_output_dtype(cfg)

    with build_accessor(cfg, image_path, kind="image", mode="test") as image_acc:
        mask_acc = build_accessor(cfg, mask_path, kind="mask", mode="test") if mask_path is not None else None
        try:
            resident = image_acc.as_tensor() if hasattr(image_acc, "as_tensor") else None
            if resident is not None and str(pad_mode) == "constant":
                vol = resident.to(dev, non_blocking=True)       # whole volume resident in HBM: gather windows on the GPU

                def read_batch(starts):
                    shifted = [tuple(s[a] - ctx[a] for a in range(3)) for s in starts]
                    return W._extract_starts(vol, shifted, read_size, "constant", cval).float()
            else:
                def read_batch(starts):                          # disk / host -> pinned -> H2D, one batch at a time
                    patches = [image_acc.read_patch(tuple(s[a] - ctx[a] for a in range(3)), read_size,
                                                    outer_pad_mode=pad_mode, outer_pad_value=cval) for s in starts]
                    host = torch.from_numpy(np.stack(patches, axis=0))
                    return host.pin_memory().to(dev, non_blocking=True) if dev.type == "cuda" else host

            def predict(batch, chunk):
                m = None
                if mask_acc is not None:
                    masks = [mask_acc.read_patch(tuple(r[0][a] - ctx[a] for a in range(3)), read_size,
                                                 outer_pad_mode="constant", outer_pad_value=0.0) for r in chunk]
                    m = torch.from_numpy(np.stack(masks, axis=0)).to(dev, non_blocking=True)
                if predictor is not None:
                    return predictor.predict(batch.float(), mask=m, mask_align_to_image=mask_align_to_image,
                                             requested_head=requested_head)
                # the reference's selection rules (utils/model_outputs.py:61-123,245-305), as its TTAPredictor applies them
                pred = pick_inference_output(cfg, forward_fn(batch), requested_head)
                if m is not None:
                    if m.shape[1] not in (1, pred.shape[1]):
                        raise ValueError(f"Mask channels {m.shape[1]} incompatible with prediction channels "
                                         f"{pred.shape[1]}")
                    pred = pred * (m > 0).to(pred.dtype)
                return pred

            out = _lazy_tile_loop(read_batch, predict, image_size=image_acc.padded_spatial_shape, roi=roi, overlap=overlap,
                                  mode=mode, region_start=region_start, region_stop=region_stop, snap_to_edge=snap,
                                  sw_batch_size=sw_batch_size, output_dtype=output_dtype, border_mask=border_mask,
                                  rank=rank, world_size=world, accumulator_reduce=accumulator_reduce, dev=dev,
                                  normalize=True, target_context=ctx,
                                  what=f" for {image_path!r}" if isinstance(image_path, (str, os.PathLike)) else "",
                                  shard_validator=shard_validator)
        finally:
            if mask_acc is not None:
                mask_acc.close()
    return out.cpu() if out.numel() else out


def lazy_predict_region(cfg, forward_fn, image_path, *, region_start: Sequence[int], region_stop: Sequence[int],
                        mask_path=None, mask_align_to_image: bool = False, device="cuda",
                        requested_head: Optional[str] = None) -> torch.Tensor:
    """``lazy.py:1261-1293`` — one bounded region (transformed/padded ZYX coordinates); windows come from the full-volume
    grid, so region boundaries see real neighbouring data."""
    return _lazy_sliding_window_cfg(cfg, forward_fn, image_path, region_start=region_start, region_stop=region_stop,
                                    mask_path=mask_path, mask_align_to_image=mask_align_to_image, device=device,
                                    requested_head=requested_head, enable_distributed_window_sharding=False)


def lazy_predict_volume(cfg, forward_fn, image_path, *, mask_path=None, mask_align_to_image: bool = False,
                        device="cuda", requested_head: Optional[str] = None) -> torch.Tensor:
    """``lazy.py:1295-1334`` — the whole volume; with ``inference.sliding_window.distributed_sharding`` inside an
    initialised process group the windows are sharded ``[rank::world]`` and the accumulators reduced onto rank 0
    (non-root ranks get an empty tensor back)."""
    from .lazy_distributed import (distributed_reduction_device, make_accumulator_reducer, should_shard_windows,
                                   validate_distributed_patch_shard)
    sc = getattr(getattr(cfg, "inference", None), "sliding_window", None)
    distributed = should_shard_windows(bool(getattr(sc, "distributed_sharding", False)))
    reducer = make_accumulator_reducer() if distributed else None
    red_dev = distributed_reduction_device(torch.device(device))

    def validator(local_count, total_count):             # lazy.py:1104-1110 -> lazy_distributed.py:110-129
        validate_distributed_patch_shard(local_count=local_count, total_count=total_count, reduction_device=red_dev)
    return _lazy_sliding_window_cfg(cfg, forward_fn, image_path, region_start=None, region_stop=None,
                                    mask_path=mask_path, mask_align_to_image=mask_align_to_image, device=device,
                                    requested_head=requested_head, enable_distributed_window_sharding=distributed,
                                    accumulator_reduce=reducer, shard_validator=validator if distributed else None)


__all__ = ["ArrayVolumeAccessor", "build_accessor", "register_accessor_factory", "get_lazy_image_reference_shape",
           "lazy_window_records", "lazy_sliding_window", "lazy_predict_region", "lazy_predict_volume"]
