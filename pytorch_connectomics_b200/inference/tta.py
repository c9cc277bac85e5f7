"""Test-time augmentation on the B200 engine — the fully-valid-channel path of
``connectomics/inference/tta.py`` / ``tta_combinations.py`` / ``tta_ensemble.py``.

Host logic (same names, argument meaning and errors as the reference):

* ``resolve_tta_augmentation_combinations(tta_cfg, spatial_dims=)`` — ``tta_combinations.py:161-193``: flip variants
  (``flip_axes``: ``"all"``/``[]``, ``"none"``/``None``, explicit lists), rotation planes (``rotation90_axes``),
  ``rotate90_k``, duplicates removed through the signature of an asymmetric probe volume;
* ``_resolve_ensemble_mode_map`` — ``tta_combinations.py:196-241`` (``"mean"|"min"|"max"`` or per-channel
  ``[[selector, mode], ...]``);
* channel selectors — ``connectomics/utils/channel_slices.py`` (ints, ``"a:b"`` strings, explicit lists).

Device work: every view is ONE gather (``pcb_tta_view`` = ``torch.flip`` + ``torch.rot90`` of ``tta.py:706-714``) and ONE
fold (``pcb_tta_fold`` = ``invert_view`` of ``tta_affinity.py:364-369`` + ``apply_preprocessing`` of ``tta.py:312-402`` +
``TTAEnsembleAccumulator._add_full_channels`` of ``tta_ensemble.py:94-110``) — un-rotating and un-flipping are index
maps inside the fold, so the per-view chain of six tensor passes becomes a single pass over the accumulator.

Not covered (loud ``NotImplementedError``): affinity-aware channel moves / partial-validity channels
(``tta_affinity.py``), ``softmax`` activations, distributed view sharding.
"""

from __future__ import annotations

import ctypes
from itertools import combinations
from typing import Any, Callable, List, Optional, Sequence, Tuple

import torch

from .. import _lib as L

_MODES = {"mean": 0, "min": 1, "max": 2}


# ----------------------------------------------------------------------------- channel selectors (utils/channel_slices.py)
def _parse_selector_string(value: str, *, context: str):
    text = value.strip()
    if not text:
        raise ValueError(f"{context} must not be empty.")
    if ":" not in text:
        try:
            return int(text)
        except ValueError as exc:
            raise ValueError(f"{context} must be an integer index or a Python-style slice string, got {value!r}.") from exc
    if text.count(":") != 1:
        raise ValueError(f"{context} must use step-free Python slice syntax 'start:end', got {value!r}.")
    a, b = text.split(":", 1)
    try:
        start = int(a.strip()) if a.strip() else None
        stop = int(b.strip()) if b.strip() else None
    except ValueError as exc:
        raise ValueError(f"{context} must use integer slice bounds in 'start:end', got {value!r}.") from exc
    return slice(start, stop)


def resolve_channel_index(index: int, *, num_channels: int, context: str = "channel selector") -> int:
    idx = int(index)
    if idx < 0:
        idx += num_channels
    if idx < 0 or idx >= num_channels:
        raise ValueError(f"Invalid {context} {index!r} for tensor with {num_channels} channels.")
    return idx


def resolve_channel_range(selector, *, num_channels: int, context: str = "channel selector") -> Tuple[int, int]:
    """``channel_slices.py::resolve_channel_range`` — contiguous selector -> absolute half-open bounds."""
    if num_channels <= 0:
        raise ValueError(f"{context} requires num_channels > 0, got {num_channels}.")
    if selector is None:
        return (0, num_channels)
    if isinstance(selector, bool) or not isinstance(selector, (int, str)):
        raise TypeError(f"{context} must be an int or a Python-style slice string, got {type(selector).__name__}.")
    parsed = selector if isinstance(selector, int) else _parse_selector_string(selector, context=context)
    if isinstance(parsed, int):
        i = resolve_channel_index(parsed, num_channels=num_channels, context=context)
        return (i, i + 1)
    start = 0 if parsed.start is None else int(parsed.start)
    stop = num_channels if parsed.stop is None else int(parsed.stop)
    if start < 0:
        start += num_channels
    if stop < 0:
        stop += num_channels
    if start < 0 or start >= num_channels:
        raise ValueError(f"Invalid {context} {selector!r} for tensor with {num_channels} channels: "
                         f"resolved start index {start} is out of bounds.")
    if stop < 0 or stop > num_channels:
        raise ValueError(f"Invalid {context} {selector!r} for tensor with {num_channels} channels: "
                         f"resolved stop index {stop} is out of bounds.")
    if stop <= start:
        raise ValueError(f"Invalid {context} {selector!r} for tensor with {num_channels} channels: "
                         f"resolved range [{start}, {stop}) is empty or inverted.")
    return (start, stop)


def resolve_channel_indices(selector, *, num_channels: int, context: str = "channel selector") -> Optional[List[int]]:
    """General selector (``None`` | int | slice string | list of ints) -> explicit channel list."""
    if selector is None:
        return None
    if isinstance(selector, (int, str)) and not isinstance(selector, bool):
        a, b = resolve_channel_range(selector, num_channels=num_channels, context=context)
        return list(range(a, b))
    if isinstance(selector, Sequence):
        if len(selector) == 0:
            raise ValueError(f"{context} must not be an empty channel list.")
        out = []
        for raw in selector:
            if isinstance(raw, str):
                try:
                    raw = int(raw.strip())
                except ValueError as exc:
                    raise ValueError(f"{context} channel lists must contain only integer indices, got {raw!r}.") from exc
            elif not isinstance(raw, int):
                raise TypeError(f"{context} channel lists must contain only integers, got {type(raw).__name__}.")
            out.append(resolve_channel_index(raw, num_channels=num_channels, context=context))
        return out
    raise TypeError(f"{context} must be an int, a slice string or a list of ints, got {type(selector).__name__}.")


# ----------------------------------------------------------------------------- augmentation combinations (tta_combinations.py)
def _to_plain_list(v) -> list:
    if isinstance(v, (list, tuple)):
        return list(v)
    if hasattr(v, "__iter__") and not isinstance(v, (str, bytes)):
        return [(_to_plain_list(e) if hasattr(e, "__iter__") and not isinstance(e, (str, bytes)) else e) for e in v]
    return [v]


def _resolve_spatial_dims(ndim: int) -> int:
    if ndim == 5:
        return 3
    if ndim == 4:
        return 2
    raise ValueError(f"Unsupported data dimensions: {ndim}")


def _normalize_spatial_axes(axes: Any, *, spatial_dims: int, context: str) -> List[int]:
    if isinstance(axes, int):
        axes = [axes]
    if not isinstance(axes, (list, tuple)):
        raise ValueError(f"{context} must be an int or list of ints, got {axes!r}.")
    out: List[int] = []
    for raw in axes:
        axis = int(raw)
        if axis < 0 or axis >= spatial_dims:
            raise ValueError(f"{context} axis must be in [0, {spatial_dims - 1}], got {axis}.")
        if axis not in out:
            out.append(axis)
    return out


def _resolve_flip_augmentations(tta_cfg, *, spatial_dims: int) -> List[List[int]]:
    cfg = getattr(tta_cfg, "flip_axes", None)
    if isinstance(cfg, str) and cfg.lower() == "none":
        return [[]]
    if cfg == "all" or cfg == []:
        out: List[List[int]] = [[]]
        for r in range(1, spatial_dims + 1):
            for combo in combinations(range(spatial_dims), r):
                out.append(list(combo))
        return out
    if cfg is None:
        return [[]]
    out = [[]]
    for raw in _to_plain_list(cfg):
        out.append(_normalize_spatial_axes(raw, spatial_dims=spatial_dims, context="flip_axes"))
    return out


def _resolve_rotation_planes(tta_cfg, *, spatial_dims: int) -> List[Tuple[int, int]]:
    cfg = getattr(tta_cfg, "rotation90_axes", None)
    if isinstance(cfg, str) and cfg.lower() == "none":
        return []
    if cfg == "all":
        if spatial_dims == 3:
            return [(0, 1), (0, 2), (1, 2)]
        if spatial_dims == 2:
            return [(0, 1)]
        raise ValueError(f"Unsupported spatial dimensions: {spatial_dims}")
    if cfg is None:
        return []
    planes: List[Tuple[int, int]] = []
    for axes in _to_plain_list(cfg):
        norm = _normalize_spatial_axes(axes, spatial_dims=spatial_dims, context="rotation90_axes")
        if len(norm) != 2:
            raise ValueError(f"Invalid rotation plane: {axes}. Each plane must contain exactly 2 axes.")
        plane = (norm[0], norm[1])
        if plane not in planes:
            planes.append(plane)
    return planes


def _resolve_rotation_k_values(tta_cfg) -> List[int]:
    cfg = getattr(tta_cfg, "rotate90_k", None)
    if cfg is None:
        return [0, 1, 2, 3]
    out: List[int] = []
    for raw in _to_plain_list(cfg):
        k = int(raw) % 4
        if k not in out:
            out.append(k)
    return out or [0]


def _augmentation_signature(*, spatial_dims: int, flip_axes, rotation_plane, k_rotations: int) -> Tuple[int, ...]:
    if spatial_dims == 3:
        base = torch.arange(2 * 3 * 5, dtype=torch.int64).reshape(2, 3, 5)
    elif spatial_dims == 2:
        base = torch.arange(2 * 5, dtype=torch.int64).reshape(2, 5)
    else:
        raise ValueError(f"Unsupported spatial dimensions: {spatial_dims}")
    if flip_axes:
        base = torch.flip(base, dims=list(flip_axes))
    if rotation_plane is not None and k_rotations % 4:
        base = torch.rot90(base, k=k_rotations, dims=rotation_plane)
    return tuple(int(v) for v in base.reshape(-1).tolist())


def resolve_tta_augmentation_combinations(tta_cfg, *, spatial_dims: int):
    """``tta_combinations.py:161-193`` — unique ``(flip_axes, rotation_plane, k)`` views (spatial axes 0=z,1=y,2=x)."""
    flips = _resolve_flip_augmentations(tta_cfg, spatial_dims=spatial_dims)
    planes = _resolve_rotation_planes(tta_cfg, spatial_dims=spatial_dims)
    if not planes:
        return [(f, None, 0) for f in flips]
    ks = _resolve_rotation_k_values(tta_cfg)
    out, seen = [], set()
    for f in flips:
        for plane in planes:
            for k in ks:
                sig = _augmentation_signature(spatial_dims=spatial_dims, flip_axes=f, rotation_plane=plane, k_rotations=k)
                if sig in seen:
                    continue
                seen.add(sig)
                out.append((f, plane, k))
    return out


def _resolve_ensemble_mode_map(ensemble_mode: Any, num_channels: int) -> List[str]:
    """``tta_combinations.py:196-241``."""
    if isinstance(ensemble_mode, str):
        return [ensemble_mode] * num_channels
    raw = _to_plain_list(ensemble_mode)
    if not isinstance(raw, list) or not raw:
        raise ValueError("ensemble_mode must be a string or a list of [channel_selector, mode] pairs, "
                         f"got {ensemble_mode!r}.")
    if isinstance(raw[0], str) and len(raw) == 1:
        return [raw[0]] * num_channels
    modes: List[Optional[str]] = [None] * num_channels
    for entry in raw:
        if not isinstance(entry, (list, tuple)) or len(entry) != 2:
            raise ValueError(f"Each ensemble_mode entry must be [channel_selector, mode], got {entry!r}.")
        selector, mode = entry
        if mode not in _MODES:
            raise ValueError(f"Unknown ensemble mode {mode!r} in per-channel spec. Use 'mean', 'min', or 'max'.")
        a, b = resolve_channel_range(str(selector), num_channels=num_channels, context="ensemble_mode channel selector")
        for ch in range(a, b):
            modes[ch] = mode
    unset = [i for i, m in enumerate(modes) if m is None]
    if unset:
        raise ValueError(f"ensemble_mode does not cover channels {unset}. Every channel must be assigned a mode.")
    return modes  # type: ignore[return-value]


# ----------------------------------------------------------------------------- activations (tta.py:141-231, 312-402)
def resolve_activation_codes(channel_activations, num_channels: int):
    """``[{channels: selector, activation: name}, ...]`` -> per-channel (code, scale): 0 none, 1 sigmoid,
    2 scale_sigmoid[:s] (default temperature 0.2), 3 tanh."""
    codes, scales = [0] * num_channels, [1.0] * num_channels
    for spec in (channel_activations or []):
        get = spec.get if isinstance(spec, dict) else (lambda k, d=None, s=spec: getattr(s, k, d))
        act = get("activation", None)
        chans = resolve_channel_indices(get("channels", None), num_channels=num_channels,
                                        context="inference.channel_activations channels")
        chans = list(range(num_channels)) if chans is None else chans
        if act is None or (isinstance(act, str) and act.lower() == "none"):
            code, scale = 0, 1.0
        elif act == "sigmoid":
            code, scale = 1, 1.0
        elif isinstance(act, str) and (act == "scale_sigmoid" or act.startswith("scale_sigmoid:")):
            code, scale = 2, 0.2
            if ":" in act:
                try:
                    scale = float(act.split(":", 1)[1])
                except ValueError as exc:
                    raise ValueError(f"Invalid scale_sigmoid scale in '{act}'. Expected 'scale_sigmoid:<float>'.") from exc
        elif act == "tanh":
            code, scale = 3, 1.0
        elif act == "softmax":
            raise NotImplementedError("pcb200 TTA: 'softmax' channel activations are not implemented in the fused fold "
                                      "kernel (sigmoid, scale_sigmoid, tanh and None are).")
        else:
            raise ValueError(f"Unknown activation '{act}' for channels {chans}. Supported: 'sigmoid', 'scale_sigmoid' "
                             "(or 'scale_sigmoid:<float>'), 'softmax', 'tanh', None")
        for c in chans:
            codes[c], scales[c] = code, scale
    return codes, scales


# ----------------------------------------------------------------------------- device ops
def _flip_mask(flip_axes) -> int:
    m = 0
    for a in (flip_axes or []):
        m |= 1 << int(a)
    return m


def apply_view(x: torch.Tensor, flip_axes, rotation_plane, k: int) -> torch.Tensor:
    """``torch.rot90(torch.flip(x, flip_axes + 2), k, rotation_plane + 2)`` for ``x: [N, C, D, H, W]`` in one gather."""
    L.require_device(x, "TTA view")
    if x.dim() != 5:
        raise ValueError(f"TTA views are implemented for 5-D tensors [N,C,D,H,W]; got shape {tuple(x.shape)}")
    x = x.contiguous()
    size = [int(v) for v in x.shape[2:]]
    ra, rb = (-1, -1) if rotation_plane is None else (int(rotation_plane[0]), int(rotation_plane[1]))
    k = int(k) % 4 if rotation_plane is not None else 0
    out_size = list(size)
    if ra >= 0 and (k & 1):
        out_size[ra], out_size[rb] = size[rb], size[ra]
    out = torch.empty((x.shape[0], x.shape[1], *out_size), device=x.device, dtype=x.dtype)
    with torch.cuda.device(x.device):
        L.check(L.lib().pcb_tta_view(L.ptr(x), L.ptr(out), L.dtype_code(x.dtype), ctypes.c_int64(int(x.shape[0] * x.shape[1])),
                                     L.i64x(size), _flip_mask(flip_axes), ra, rb, k, L.stream_ptr(x.device)), "pcb_tta_view")
    return out


class TTAEnsemble:
    """Streaming ensemble over augmentation views (``tta.py:691-771`` + ``tta_ensemble.py``, full channels).

    ``predict(images, network_fn)``: for every view ``x_aug = view(images)``, ``pred = network_fn(x_aug)`` (a model call
    or a sliding-window engine call), then one fold kernel un-views ``pred``, applies the channel activations and the
    channel selection, casts to ``output_dtype`` and updates the running mean / min / max."""

    def __init__(self, tta_cfg=None, *, channel_activations=None, select_channel=None,
                 output_dtype: Optional[torch.dtype] = None) -> None:
        self.tta_cfg = tta_cfg
        self.channel_activations = channel_activations
        self.select_channel = select_channel
        self.output_dtype = output_dtype
        if tta_cfg is not None and getattr(tta_cfg, "affinity_offsets", None):
            raise NotImplementedError("pcb200 TTA: affinity-aware channel moves (tta_affinity.py) are not implemented")

    def combinations(self, ndim: int):
        if self.tta_cfg is None or not getattr(self.tta_cfg, "enabled", True):
            return [([], None, 0)]
        return resolve_tta_augmentation_combinations(self.tta_cfg, spatial_dims=_resolve_spatial_dims(ndim))

    def predict(self, images: torch.Tensor, network_fn: Callable[[torch.Tensor], torch.Tensor]) -> torch.Tensor:
        if images.dim() != 5:
            raise ValueError(f"pcb200 TTA expects [N,C,D,H,W] inputs; got shape {tuple(images.shape)}")
        combos = self.combinations(images.dim())
        mode_cfg = getattr(self.tta_cfg, "ensemble_mode", "mean") if self.tta_cfg is not None else "mean"
        acc = None
        src = modes = acts = scales = None
        size = [int(v) for v in images.shape[2:]]
        for n_prev, (flip_axes, plane, k) in enumerate(combos):
            trivial = not flip_axes and (plane is None or k % 4 == 0)
            x_aug = images if trivial else apply_view(images, flip_axes, plane, k)
            pred = network_fn(x_aug)
            if not isinstance(pred, torch.Tensor) or pred.dim() != 5:
                raise ValueError("pcb200 TTA: `network_fn` must return a [N,C,D,H,W] tensor")
            L.require_device(pred, "TTA fold")
            pred = pred.contiguous()
            c_pred = int(pred.shape[1])
            if acc is None:
                sel = resolve_channel_indices(self.select_channel, num_channels=c_pred,
                                              context="inference.model.select_channel")
                src = list(range(c_pred)) if sel is None else sel
                codes, sc = resolve_activation_codes(self.channel_activations, c_pred)
                acts, scales = [codes[c] for c in src], [sc[c] for c in src]
                modes = [_MODES.get(m, -1) for m in _resolve_ensemble_mode_map(mode_cfg, len(src))]
                if any(m < 0 for m in modes):
                    raise ValueError(f"Unknown TTA ensemble modes: {sorted(set(_resolve_ensemble_mode_map(mode_cfg, len(src))) - set(_MODES))}.")
                odt = self.output_dtype or pred.dtype
                acc = torch.empty((pred.shape[0], len(src), *size), device=pred.device, dtype=odt)
            ra, rb = (-1, -1) if plane is None else (int(plane[0]), int(plane[1]))
            with torch.cuda.device(pred.device):
                L.check(L.lib().pcb_tta_fold(
                    L.ptr(pred), L.dtype_code(pred.dtype), L.ptr(acc), L.dtype_code(acc.dtype), ctypes.c_int64(int(pred.shape[0])),
                    ctypes.c_int64(c_pred), ctypes.c_int64(len(src)), L.i64x(size), _flip_mask(flip_axes), ra, rb,
                    (int(k) % 4) if plane is not None else 0, (ctypes.c_int * len(src))(*src), (ctypes.c_int * len(src))(*modes),
                    (ctypes.c_int * len(src))(*acts), (ctypes.c_float * len(src))(*scales), n_prev, L.stream_ptr(pred.device)),
                    "pcb_tta_fold")
        return acc


__all__ = ["TTAEnsemble", "apply_view", "resolve_activation_codes", "resolve_channel_indices", "resolve_channel_range",
           "resolve_tta_augmentation_combinations", "_resolve_ensemble_mode_map"]
