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

Affinity-aware inversion (``tta_affinity.py``: channel moves + roll shifts with invalid wrapped faces), the validity-aware
aggregation of partial channels (``tta_ensemble.py:121-211``), ``softmax`` activations, distributed view sharding
(``tta.py:771-804,1341-1519``) and the patch-first local loop (``tta.py:880-1314``) are built on the same two kernels
(``pcb_tta_fold_ex`` / ``pcb_tta_unview``) — see ``TTAEnsembleAccumulator``, ``TTAEnsemble.predict`` and
``TTAEnsemble.predict_patch_first``.
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


def _canonical_range_selector(selector, *, context: str):
    """``normalize_channel_range_selector`` (channel_slices.py:55-81): None | int | the selector string re-rendered as
    ``"start:stop"`` without blanks (error messages quote THIS form, not the raw string)."""
    if selector is None:
        return None
    if isinstance(selector, int):
        return int(selector)
    if isinstance(selector, str):
        parsed = _parse_selector_string(selector, context=context)
        if isinstance(parsed, int):
            return parsed
        return f"{'' if parsed.start is None else parsed.start}:{'' if parsed.stop is None else parsed.stop}"
    raise TypeError(f"{context} must be an int or a Python-style slice string, got {type(selector).__name__}.")


def resolve_channel_index(index: int, *, num_channels: int, context: str = "channel selector") -> int:
    """``channel_slices.py:130-145`` — Python indexing rules for one (possibly negative) channel index."""
    idx = int(index)
    if idx < 0:
        idx += num_channels
    if not 0 <= idx < num_channels:
        raise ValueError(f"Invalid {context} {index!r} for tensor with {num_channels} channels: "
                         f"resolved index {idx} is out of bounds.")
    return idx


def resolve_channel_range(selector, *, num_channels: int, context: str = "channel selector") -> Tuple[int, int]:
    """``channel_slices.py:148-193`` — contiguous selector -> absolute half-open bounds."""
    if num_channels <= 0:
        raise ValueError(f"{context} requires num_channels > 0, got {num_channels}.")
    canon = _canonical_range_selector(selector, context=context)
    if canon is None:
        return (0, num_channels)
    if isinstance(canon, int):
        i = resolve_channel_index(canon, num_channels=num_channels, context=context)
        return (i, i + 1)
    lo_text, hi_text = canon.split(":", 1)
    start = int(lo_text) if lo_text else 0
    stop = int(hi_text) if hi_text else num_channels
    if start < 0:
        start += num_channels
    if stop < 0:
        stop += num_channels
    where = f"Invalid {context} {canon!r} for tensor with {num_channels} channels: "
    if not 0 <= start < num_channels:
        raise ValueError(where + f"resolved start index {start} is out of bounds.")
    if not 0 <= stop <= num_channels:
        raise ValueError(where + f"resolved stop index {stop} is out of bounds.")
    if stop <= start:
        raise ValueError(where + f"resolved range [{start}, {stop}) is empty or inverted.")
    return (start, stop)


def resolve_channel_indices(selector, *, num_channels: int, context: str = "channel selector") -> List[int]:
    """``channel_slices.py:196-224`` — general selector (``None`` = every channel | int | slice string | list of ints) ->
    explicit channel list."""
    if num_channels <= 0:
        raise ValueError(f"{context} requires num_channels > 0, got {num_channels}.")
    if selector is None:
        return list(range(num_channels))
    if isinstance(selector, (int, str)):
        a, b = resolve_channel_range(selector, num_channels=num_channels, context=context)
        return list(range(a, b))
    if isinstance(selector, Sequence) and not isinstance(selector, bytes):
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
            out.append(raw)
        return [resolve_channel_index(v, num_channels=num_channels, context=context) for v in out]
    raise TypeError(f"{context} must be an int, a Python-style slice string, or an explicit list of ints; "
                    f"got {type(selector).__name__}.")


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
def resolve_activation_specs(channel_activations, num_channels: int):
    """``[{channels: selector, activation: name}, ...]`` -> per-channel (code, scale, softmax group): 0 none, 1 sigmoid,
    2 scale_sigmoid[:s] (default temperature 0.2), 3 tanh, 4 softmax over ``group`` (the spec's channel list; a
    single-channel softmax is skipped like the reference does, ``tta.py:368-375``).  The reference applies the specs one
    after the other IN PLACE; a channel named by two active specs would get both — that composition is refused here."""
    codes, scales = [0] * num_channels, [1.0] * num_channels
    groups: List[Optional[List[int]]] = [None] * num_channels
    seen = set()
    for spec in (channel_activations or []):
        get = spec.get if isinstance(spec, dict) else (lambda k, d=None, s=spec: getattr(s, k, d))
        act = get("activation", None)
        chans = resolve_channel_indices(get("channels", None), num_channels=num_channels,
                                        context="inference.channel_activations channels")
        chans = list(range(num_channels)) if chans is None else chans
        group = None
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
            code, scale = (4, 1.0) if len(chans) > 1 else (0, 1.0)
            group = list(chans) if len(chans) > 1 else None
        else:
            raise ValueError(f"Unknown activation '{act}' for channels {chans}. Supported: 'sigmoid', 'scale_sigmoid' "
                             "(or 'scale_sigmoid:<float>'), 'softmax', 'tanh', None")
        for c in chans:
            if code != 0:
                if c in seen:
                    raise NotImplementedError(f"pcb200 TTA: channel {c} is named by two channel_activations entries; "
                                              "composed activations are not implemented")
                seen.add(c)
            codes[c], scales[c], groups[c] = code, scale, group
    return codes, scales, groups


def resolve_activation_codes(channel_activations, num_channels: int):
    codes, scales, _groups = resolve_activation_specs(channel_activations, num_channels)
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


def _int_arr(vals):
    vals = [int(v) for v in vals]
    return (ctypes.c_int * max(1, len(vals)))(*vals) if vals else None


class TTAEnsembleAccumulator:
    """``tta_ensemble.py:13-211`` on the device: fully valid channels through the running mean / min / max, partial
    (affinity, re-anchored) channels through fp32 statistics + per-voxel contribution counts, both updated by ONE kernel
    per view.  ``add(prediction, validity)`` is the reference's call (canonical, pre-processed prediction);
    ``fold_view`` is the fused one: it takes the RAW view-frame network output and does the inversion, channel moves, roll
    shifts, activations, channel selection and the update in the same pass."""

    def __init__(self, shape, *, dtype: torch.dtype, device, mode_map: Sequence[str], partial_channels: Sequence[int],
                 distributed_sharding: bool, max_views: int) -> None:
        self.shape = tuple(int(v) for v in shape)
        self.dtype = dtype
        self.device = torch.device(device)
        self.mode_map = tuple(str(m) for m in mode_map)
        if len(self.shape) != 5 or len(self.mode_map) != self.shape[1]:
            raise ValueError(f"Invalid TTA accumulator shape/modes: shape={self.shape}, modes={len(self.mode_map)}.")
        bad = sorted(set(self.mode_map) - set(_MODES))
        if bad:
            raise ValueError(f"Unknown TTA ensemble modes: {bad}.")
        self.partial_channels = tuple(sorted({int(c) for c in partial_channels}))
        if any(c < 0 or c >= self.shape[1] for c in self.partial_channels):
            raise ValueError(f"Partial TTA channels {self.partial_channels} are invalid for {self.shape[1]} "
                             "output channels.")
        pset = set(self.partial_channels)
        self.full_channels = tuple(c for c in range(self.shape[1]) if c not in pset)
        self.distributed_sharding = bool(distributed_sharding)
        self.num_predictions = 0
        if self.device.type != "cuda":
            raise RuntimeError("pcb200 TTAEnsembleAccumulator runs on a CUDA device only (no CPU fallback)")
        self.legacy_result = torch.zeros(self.shape, device=self.device, dtype=dtype)
        pshape = (self.shape[0], len(self.partial_channels), *self.shape[2:])
        self.partial_statistics = torch.empty(pshape, device=self.device, dtype=torch.float32)
        for j, c in enumerate(self.partial_channels):
            m = self.mode_map[c]
            self.partial_statistics[:, j].fill_(0.0 if m == "mean" else (float("inf") if m == "min" else float("-inf")))
        self.count_dtype = torch.uint8 if int(max_views) < 256 else torch.int16
        self.partial_counts = torch.zeros(pshape, device=self.device, dtype=self.count_dtype)

    @property
    def has_partial_channels(self) -> bool:
        return bool(self.partial_channels)

    # one launch: every accumulator channel from prediction channel src[c] (+ shift), activation, aggregate
    def _launch(self, pred: torch.Tensor, *, flip_axes, plane, k, src, shift, acts, scales, sm_off, sm_len, sm_src, sm_shift,
                vmask: Optional[torch.Tensor]) -> None:
        L.require_device(pred, "TTA fold")
        pred = pred.contiguous()
        cacc = self.shape[1]
        size = [int(v) for v in self.shape[2:]]
        ra, rb = (-1, -1) if plane is None else (int(plane[0]), int(plane[1]))
        kk = (int(k) % 4) if plane is not None else 0
        vsize = list(size)
        if ra >= 0 and (kk & 1):
            vsize[ra], vsize[rb] = size[rb], size[ra]
        if tuple(int(v) for v in pred.shape[2:]) != tuple(vsize) or int(pred.shape[0]) != self.shape[0]:
            raise ValueError(f"TTA prediction shape {tuple(pred.shape)} does not match accumulator shape {self.shape} "
                             f"for view (flip={list(flip_axes or [])}, plane={plane}, k={k}).")
        modes = [_MODES[m] for m in self.mode_map]
        part = [-1] * cacc
        for j, c in enumerate(self.partial_channels):
            part[c] = j
        cpart = len(self.partial_channels)
        flat_shift = [int(v) for sh in shift for v in sh]
        flat_sm_shift = [int(v) for sh in sm_shift for v in sh]
        with torch.cuda.device(pred.device):
            L.check(L.lib().pcb_tta_fold_ex(
                L.ptr(pred), L.dtype_code(pred.dtype), L.ptr(self.legacy_result), L.dtype_code(self.dtype),
                ctypes.c_int64(self.shape[0]), ctypes.c_int64(int(pred.shape[1])), ctypes.c_int64(cacc), L.i64x(size),
                _flip_mask(flip_axes), ra, rb, kk, _int_arr(src), _int_arr(modes), _int_arr(acts),
                (ctypes.c_float * cacc)(*[float(v) for v in scales]), _int_arr(flat_shift), _int_arr(sm_off), _int_arr(sm_len),
                _int_arr(sm_src), _int_arr(flat_sm_shift), len(sm_src), _int_arr(part), _int_arr(modes), ctypes.c_int64(cpart),
                L.ptr(self.partial_statistics) if cpart else None, L.ptr(self.partial_counts) if cpart else None,
                0 if self.count_dtype == torch.uint8 else 1, L.ptr(vmask) if vmask is not None else None,
                1 if self.distributed_sharding else 0, int(self.num_predictions), L.stream_ptr(pred.device)), "pcb_tta_fold_ex")
        self.num_predictions += 1

    def add(self, prediction: torch.Tensor, validity) -> None:
        """Stream one pre-processed canonical prediction (``tta_ensemble.py:164-185``)."""
        if tuple(prediction.shape) != self.shape:
            raise ValueError(f"TTA prediction shape {tuple(prediction.shape)} does not match accumulator "
                             f"shape {self.shape}.")
        if len(validity.channels) != self.shape[1]:
            raise ValueError(f"TTA validity describes {len(validity.channels)} channels, expected {self.shape[1]}.")
        cacc = self.shape[1]
        vmask = None
        entries = [validity.channels[c] for c in self.partial_channels]
        if any(e is not None for e in entries):
            vol = self.shape[2:]
            vmask = torch.ones((len(entries), *vol), device=self.device, dtype=torch.uint8)
            for j, e in enumerate(entries):
                if e is None:
                    continue
                if isinstance(e, tuple):
                    m = torch.zeros(vol, device=self.device, dtype=torch.uint8)
                    m[e] = 1
                else:
                    m = e.to(device=self.device, dtype=torch.bool)
                    if m.dim() == len(vol) + 1:
                        if m.shape[0] != 1:
                            raise NotImplementedError("pcb200 TTA: per-sample validity masks are not implemented")
                        m = m[0]
                    if tuple(m.shape) != tuple(vol):
                        raise ValueError(f"TTA validity shape {tuple(m.shape)} does not match channel value shape "
                                         f"{(self.shape[0], *vol)}.")
                    m = m.to(torch.uint8)
                vmask[j] = m
        incoming = prediction if prediction.dtype == self.dtype else prediction.to(self.dtype)
        self._launch(incoming, flip_axes=[], plane=None, k=0, src=list(range(cacc)), shift=[(0, 0, 0)] * cacc,
                     acts=[0] * cacc, scales=[1.0] * cacc, sm_off=[0] * cacc, sm_len=[0] * cacc, sm_src=[], sm_shift=[], vmask=vmask)

    def fold_view(self, pred: torch.Tensor, *, flip_axes, plane, k, view_plan, selection: Sequence[int], act_specs) -> None:
        """Fused path.  ``pred`` is the network output in the VIEW frame (raw channels); ``selection`` the raw canonical channel
        of every accumulator channel; ``act_specs`` = ``resolve_activation_specs`` over the raw channels."""
        from .tta_affinity import view_channel_maps
        craw = int(pred.shape[1])
        rsrc, rshift = view_channel_maps(view_plan, craw)
        codes, scales, groups = act_specs
        src = [rsrc[r] for r in selection]
        shift = [rshift[r] for r in selection]
        acts = [codes[r] for r in selection]
        scl = [scales[r] for r in selection]
        sm_off, sm_len, sm_src, sm_shift = [0] * len(selection), [0] * len(selection), [], []
        seen = {}
        for c, r in enumerate(selection):
            if acts[c] == 4:
                key = tuple(groups[r])
                if key not in seen:
                    seen[key] = (len(sm_src), len(key))
                    sm_src += [rsrc[m] for m in key]
                    sm_shift += [rshift[m] for m in key]
                sm_off[c], sm_len[c] = seen[key]
        self._launch(pred, flip_axes=flip_axes, plane=plane, k=k, src=src, shift=shift, acts=acts, scales=scl, sm_off=sm_off,
                     sm_len=sm_len, sm_src=sm_src, sm_shift=sm_shift, vmask=None)

    def finalize(self, *, legacy_result=None, partial_statistics=None, partial_counts=None) -> torch.Tensor:
        """``tta_ensemble.py:187-211``: the aggregate; a partial channel without any valid contribution is an error."""
        result = (self.legacy_result if legacy_result is None else legacy_result).clone()
        if not self.partial_channels:
            return result
        stats = self.partial_statistics if partial_statistics is None else partial_statistics
        counts = self.partial_counts if partial_counts is None else partial_counts
        if counts.dtype not in (torch.uint8, torch.int16):
            counts = counts.to(torch.int16)
        stats, counts = stats.contiguous(), counts.contiguous()
        flag = torch.full((1,), -1, device=self.device, dtype=torch.int64)       # all ones = UINT64_MAX
        nvox = 1
        for v in self.shape[2:]:
            nvox *= v
        pm = [_MODES[self.mode_map[c]] for c in self.partial_channels]
        with torch.cuda.device(self.device):
            L.check(L.lib().pcb_tta_finalize_partial(
                L.ptr(stats), L.ptr(counts), 0 if counts.dtype == torch.uint8 else 1, L.ptr(result), L.dtype_code(result.dtype),
                ctypes.c_int64(self.shape[0]), ctypes.c_int64(self.shape[1]), ctypes.c_int64(len(pm)), _int_arr(self.partial_channels),
                _int_arr(pm), ctypes.c_int64(nvox), L.ptr(flag), L.stream_ptr(self.device)), "pcb_tta_finalize_partial")
        first = int(flag.item())
        if first != -1:
            j = (first // nvox) % len(pm)
            rem = first % nvox
            idx = [first // (nvox * len(pm))]
            for d in reversed(self.shape[2:]):
                idx.insert(1, rem % d)
                rem //= d
            raise RuntimeError(f"TTA ensemble has zero valid contributions for channel {self.partial_channels[j]} at "
                               f"voxel index {tuple(idx)}.")
        return result

    def reduce_to_rank_zero(self, group=None) -> Optional[torch.Tensor]:
        """Distributed view sharding (``tta.py:1341-1519``): SUM / MIN / MAX of the per-rank accumulators into rank 0, "mean"
        channels divided by the total number of views; other ranks get ``None``."""
        import torch.distributed as dist
        rank = dist.get_rank(group)
        ops = {"mean": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX}
        n = torch.tensor([self.num_predictions], device=self.device, dtype=torch.int64)
        dist.reduce(n, dst=0, op=dist.ReduceOp.SUM, group=group)
        legacy = self.legacy_result.clone()
        for mode in sorted(set(self.mode_map)):
            red = self.legacy_result.clone()
            dist.reduce(red, dst=0, op=ops[mode], group=group)
            for c in self.full_channels:
                if self.mode_map[c] == mode:
                    legacy[:, c] = red[:, c] / float(int(n)) if (mode == "mean" and rank == 0) else red[:, c]
        stats = counts = None
        if self.partial_channels:
            stats = self.partial_statistics.clone()
            for mode in sorted({self.mode_map[c] for c in self.partial_channels}):
                red = self.partial_statistics.clone()
                dist.reduce(red, dst=0, op=ops[mode], group=group)
                for j, c in enumerate(self.partial_channels):
                    if self.mode_map[c] == mode:
                        stats[:, j] = red[:, j]
            counts = self.partial_counts.to(torch.int32)
            dist.reduce(counts, dst=0, op=dist.ReduceOp.SUM, group=group)
        if rank != 0:
            return None
        if int(n) <= 0:
            raise RuntimeError("Distributed TTA sharding reduced zero predictions on rank 0.")
        return self.finalize(legacy_result=legacy, partial_statistics=stats, partial_counts=counts)


class TTAEnsemble:
    """Streaming ensemble over augmentation views (``tta.py:691-878`` volume-first, ``:880-1314`` patch-first local).

    ``predict(images, network_fn)``: for every view ``x_aug = view(images)``, ``pred = network_fn(x_aug)`` (a model call
    or a sliding-window engine call), then one fold kernel un-views ``pred`` (with the affinity channel moves / roll shifts
    of ``tta_affinity.py`` when a plan applies), applies the channel activations and the channel selection, casts to
    ``output_dtype`` and updates the running mean / min / max (fully valid channels) or the validity-aware statistics
    (partial channels).  ``distributed_sharding=True`` gives every rank the views ``rank::world`` and reduces to rank 0."""

    def __init__(self, tta_cfg=None, *, channel_activations=None, select_channel=None,
                 output_dtype: Optional[torch.dtype] = None, cfg=None, affinity_groups=None, affinity_mode: Optional[str] = None,
                 requested_head: Optional[str] = None, distributed_sharding: bool = False, process_group=None) -> None:
        self.tta_cfg = tta_cfg
        self.channel_activations = channel_activations
        self.select_channel = select_channel
        self.output_dtype = output_dtype
        self.cfg = cfg
        self.affinity_groups = affinity_groups
        self.affinity_mode = affinity_mode
        self.requested_head = requested_head
        self.distributed_sharding = bool(distributed_sharding)
        self.process_group = process_group

    def combinations(self, ndim: int):
        if self.tta_cfg is None or not getattr(self.tta_cfg, "enabled", True):
            return [([], None, 0)]
        return resolve_tta_augmentation_combinations(self.tta_cfg, spatial_dims=_resolve_spatial_dims(ndim))

    # ---- shared set-up -------------------------------------------------------------------------------------------
    def _local_indices(self, n_views: int) -> List[int]:
        idx = list(range(n_views))
        if not self.distributed_sharding:
            return idx
        import torch.distributed as dist
        rank, world = dist.get_rank(self.process_group), dist.get_world_size(self.process_group)
        idx = idx[rank::world]
        if not idx:
            raise RuntimeError("Distributed TTA sharding produced an empty augmentation shard for "
                               f"rank {rank}. Reduce the GPU count or increase TTA variants.")
        return idx

    def _affinity_plan(self, combos, num_raw: int):
        from .tta_affinity import build_affinity_tta_plan
        if self.affinity_groups is not None:
            return build_affinity_tta_plan(None, augmentation_combinations=combos, num_raw=num_raw, groups=self.affinity_groups,
                                           mode=self.affinity_mode)
        if self.cfg is not None:
            return build_affinity_tta_plan(self.cfg, augmentation_combinations=combos, num_raw=num_raw,
                                           requested_head=self.requested_head)
        return None

    def _make_accumulator(self, pred: torch.Tensor, size, combos, plan):
        c_raw = int(pred.shape[1])
        sel = resolve_channel_indices(self.select_channel, num_channels=c_raw, context="inference.model.select_channel")
        selection = list(range(c_raw)) if sel is None else sel
        specs = resolve_activation_specs(self.channel_activations, c_raw)
        mode_cfg = getattr(self.tta_cfg, "ensemble_mode", "mean") if self.tta_cfg is not None else "mean"
        mode_map = _resolve_ensemble_mode_map(mode_cfg, len(selection))
        bad = sorted(set(mode_map) - set(_MODES))
        if bad:
            raise ValueError(f"Unknown TTA ensemble modes: {bad}.")
        partial = [] if plan is None else [i for i, r in enumerate(selection) if r in plan.partial_channels]
        acc = TTAEnsembleAccumulator((int(pred.shape[0]), len(selection), *size), dtype=self.output_dtype or pred.dtype,
                                     device=pred.device, mode_map=mode_map, partial_channels=partial,
                                     distributed_sharding=self.distributed_sharding, max_views=len(combos))
        return acc, selection, specs

    def _finish(self, acc: TTAEnsembleAccumulator) -> torch.Tensor:
        if self.distributed_sharding:
            out = acc.reduce_to_rank_zero(self.process_group)
            return torch.empty(0, device=acc.device) if out is None else out
        return acc.finalize()

    # ---- volume-first (tta.py:691-878) -----------------------------------------------------------------------------
    def predict(self, images: torch.Tensor, network_fn: Callable[[torch.Tensor], torch.Tensor]) -> torch.Tensor:
        if images.dim() != 5:
            raise ValueError(f"pcb200 TTA expects [N,C,D,H,W] inputs; got shape {tuple(images.shape)}")
        combos = self.combinations(images.dim())
        size = [int(v) for v in images.shape[2:]]
        acc = plan = selection = specs = None
        for vi in self._local_indices(len(combos)):
            flip_axes, plane, k = combos[vi]
            trivial = not flip_axes and (plane is None or k % 4 == 0)
            x_aug = images if trivial else apply_view(images, flip_axes, plane, k)
            pred = network_fn(x_aug)
            if not isinstance(pred, torch.Tensor) or pred.dim() != 5:
                raise ValueError("pcb200 TTA: `network_fn` must return a [N,C,D,H,W] tensor")
            L.require_device(pred, "TTA fold")
            if acc is None:
                plan = self._affinity_plan(combos, int(pred.shape[1]))
                acc, selection, specs = self._make_accumulator(pred, size, combos, plan)
            from .tta_affinity import validate_affinity_output
            validate_affinity_output(plan, pred)
            acc.fold_view(pred, flip_axes=flip_axes, plane=plane, k=k, view_plan=None if plan is None else plan.views[vi],
                          selection=selection, act_specs=specs)
        return self._finish(acc)

    # ---- patch-first local (tta.py:880-1314) -------------------------------------------------------------------
    def predict_patch_first(self, images: torch.Tensor, network_fn: Callable[[torch.Tensor], torch.Tensor], *, roi_size,
                            overlap=0.5, sw_batch_size: int = 1, mode: str = "constant", padding_mode: str = "constant",
                            cval: float = 0.0) -> torch.Tensor:
        """Slide ONCE over the volume and evaluate every local view inside each window batch: crop -> view -> network ->
        inverse view (+ affinity moves) -> per-view overlap-add accumulators; per-view normalisation, activations and the
        ensemble follow at the end, exactly in the reference's order.  Fully valid channels accumulate in the output dtype
        against one fp32 weight volume; partial channels in fp32 against one weight volume per roll shift, restricted to the
        shift's valid box, and voxels without coverage are invalid for the ensemble."""
        from . import window as W
        from .tta_affinity import ViewValidity, invert_view, valid_slices_for_shift
        if images.dim() != 5:
            raise ValueError(f"pcb200 TTA expects [N,C,D,H,W] inputs; got shape {tuple(images.shape)}")
        L.require_device(images, "patch-first TTA")
        combos = self.combinations(images.dim())
        roi = tuple(int(v) for v in roi_size)
        local = self._local_indices(len(combos))
        odt = self.output_dtype
        outputs = []
        for b in range(int(images.shape[0])):
            sample = images[b:b + 1]
            original = tuple(int(v) for v in sample.shape[2:])
            self._validate_patch_first_local_supported([combos[i] for i in local], image_size=original, roi_size=roi)
            padded = tuple(max(original[a], roi[a]) for a in range(3))
            ov = overlap if isinstance(overlap, (tuple, list)) else (float(overlap),) * 3
            interval = W.compute_scan_interval(padded, roi, 3, ov)
            slices = W.dense_patch_slices(padded, roi, interval, return_slice=True)
            full_acc: List[Optional[torch.Tensor]] = [None] * len(local)
            part_acc: List[Optional[torch.Tensor]] = [None] * len(local)
            plan = None
            n_raw = None
            raw_full: List[int] = []
            raw_part: List[int] = []
            w_full = None
            w_part = {}
            scratch_w = {}
            value_map = value_map32 = pmap32 = ones_patch = None
            dev = sample.device
            for s0 in range(0, len(slices), int(sw_batch_size)):
                cur = slices[s0:s0 + int(sw_batch_size)]
                batch, locs = W._extract_padded_patch_batch(sample, cur, roi_size=roi, padding_mode=padding_mode, cval=cval)
                batch = batch.float()
                weights_added = False
                for li, vi in enumerate(local):
                    flip_axes, plane, k = combos[vi]
                    trivial = not flip_axes and (plane is None or k % 4 == 0)
                    x_aug = batch if trivial else apply_view(batch, flip_axes, plane, k)
                    pred = network_fn(x_aug)
                    if not isinstance(pred, torch.Tensor) or pred.dim() != 5:
                        raise ValueError("pcb200 TTA: `network_fn` must return a [N,C,D,H,W] tensor")
                    if n_raw is None:
                        n_raw = int(pred.shape[1])
                        plan = self._affinity_plan(combos, n_raw)
                        pset = set() if plan is None else set(plan.partial_channels)
                        raw_part = sorted(pset)
                        raw_full = [c for c in range(n_raw) if c not in pset]
                        if odt is None:
                            odt = pred.dtype
                        value_map, _ = W.build_sliding_accumulator_weight_maps(roi, mode=mode, device=dev, value_dtype=odt)
                        value_map32 = value_map.float()
                        pmap32, _ = W.build_sliding_accumulator_weight_maps(roi, mode=mode, device=dev, value_dtype=torch.float32)
                        ones_patch = torch.ones((1, 1, *roi), device=dev, dtype=torch.float32)
                        if raw_full:
                            w_full = torch.zeros((1, 1, *padded), device=dev, dtype=torch.float32)
                        if raw_part:
                            keys = {()} | (set(plan.shifts) if plan is not None else set())
                            w_part = {key: torch.zeros((1, 1, *padded), device=dev, dtype=torch.float32) for key in keys}
                    elif int(pred.shape[1]) != n_raw:
                        raise RuntimeError("Patch-first local TTA model output channel count changed between "
                                           f"views: expected {n_raw}, got {int(pred.shape[1])}.")
                    view_plan = None if plan is None else plan.views[vi]
                    pred, _validity = invert_view(pred, flip_axes=flip_axes, rotation_plane_spatial=plane, k=k,
                                                  view_plan=view_plan, tta_plan=plan)
                    if tuple(int(v) for v in pred.shape[2:]) != roi:
                        raise RuntimeError("Patch-first local TTA requires patch predictions to preserve "
                                           f"the ROI spatial shape. Got prediction.shape={tuple(pred.shape)} "
                                           f"and roi_size={roi}.")
                    if not weights_added:
                        # weight volumes: sum of the map over the covering windows, in fp32 (value 1.0 * map is exact)
                        if w_full is not None:
                            sw = scratch_w.setdefault("f", torch.zeros_like(w_full))
                            for loc in locs:
                                W._accumulate_window(ones_patch, value_map32, w_full, sw, roi, padded, (0, 0, 0), loc, roi)
                        for key, wacc in w_part.items():
                            box = tuple(slice(0, r) for r in roi) if not key else valid_slices_for_shift(roi, key)
                            lo = tuple(int(sl.start) for sl in box)
                            ext = tuple(int(sl.stop) - int(sl.start) for sl in box)
                            if min(ext) <= 0:
                                continue
                            sw = scratch_w.setdefault("p", torch.zeros_like(wacc))
                            for loc in locs:
                                W._accumulate_window(ones_patch, pmap32, wacc, sw, roi, padded, lo,
                                                     tuple(loc[a] + lo[a] for a in range(3)), ext)
                        weights_added = True
                    if raw_full:
                        if full_acc[li] is None:
                            full_acc[li] = torch.zeros((1, len(raw_full), *padded), device=dev, dtype=odt)
                        pf = pred if len(raw_full) == n_raw else pred[:, raw_full]
                        pf = pf.to(odt).contiguous()
                        sw = scratch_w.setdefault("fv", torch.zeros((1, 1, *padded), device=dev, dtype=odt))
                        W._accumulate_batch(pf, value_map, full_acc[li], sw, roi, padded, locs)
                    if raw_part:
                        if part_acc[li] is None:
                            part_acc[li] = torch.zeros((1, len(raw_part), *padded), device=dev, dtype=torch.float32)
                        pp = pred[:, raw_part].float().contiguous()
                        sw = scratch_w.setdefault("pv", torch.zeros((1, 1, *padded), device=dev, dtype=torch.float32))
                        W._accumulate_batch(pp, pmap32, part_acc[li], sw, roi, padded, locs)
            if n_raw is None:
                raise RuntimeError("Patch-first local TTA generated no predictions.")
            crop = tuple(slice(0, v) for v in original)
            acc = None
            sel = resolve_channel_indices(self.select_channel, num_channels=n_raw, context="inference.model.select_channel")
            selection = list(range(n_raw)) if sel is None else sel
            specs = resolve_activation_specs(self.channel_activations, n_raw)
            for li, vi in enumerate(local):
                raw_volume = torch.zeros((1, n_raw, *padded), device=dev, dtype=odt)
                raw_validity: list = [None] * n_raw
                if raw_full:
                    raw_volume[:, raw_full] = W.normalize_weighted_accumulator(full_acc[li], w_full.clone())
                    full_acc[li] = None
                if raw_part:
                    view_plan = None if plan is None else plan.views[vi]
                    for pj, rc in enumerate(raw_part):
                        sh = None if view_plan is None else view_plan.shift_for_channel(rc)
                        weight = w_part[() if sh is None else tuple(sh)][:, 0]
                        cov = weight > 0
                        norm = torch.zeros_like(weight)
                        norm[cov] = part_acc[li][:, pj][cov] / weight[cov]
                        raw_volume[:, rc] = norm.to(odt)
                        raw_validity[rc] = cov
                    part_acc[li] = None
                raw_volume = raw_volume[(slice(None), slice(None), *crop)].contiguous()
                validity = ViewValidity(tuple(v[(slice(None), *crop)] if torch.is_tensor(v) else v for v in raw_validity)).select(selection)
                if acc is None:
                    acc, selection, specs = self._make_accumulator(raw_volume, list(original), combos, plan)
                # pre-processing (activations + selection + cast) of the canonical volume is the fold's own first half; the
                # tensor validity goes in as a mask: identity geometry, no channel moves (they were applied per window)
                self._fold_canonical(acc, raw_volume, selection, specs, validity)
            outputs.append(self._finish(acc))
        if any(o.numel() == 0 for o in outputs):
            return torch.empty(0, device=images.device)
        return torch.cat(outputs, dim=0)

    def _fold_canonical(self, acc: TTAEnsembleAccumulator, raw_volume: torch.Tensor, selection, specs, validity) -> None:
        codes, scales, groups = specs
        vmask = None
        entries = [validity.channels[c] for c in acc.partial_channels]
        if any(e is not None for e in entries):
            vol = acc.shape[2:]
            vmask = torch.ones((len(entries), *vol), device=acc.device, dtype=torch.uint8)
            for j, e in enumerate(entries):
                if e is not None:
                    vmask[j] = (e[0] if e.dim() == len(vol) + 1 else e).to(torch.uint8)
        sm_off, sm_len, sm_src, sm_shift = [0] * len(selection), [0] * len(selection), [], []
        seen = {}
        for c, r in enumerate(selection):
            if codes[r] == 4:
                key = tuple(groups[r])
                if key not in seen:
                    seen[key] = (len(sm_src), len(key))
                    sm_src += list(key)
                    sm_shift += [(0, 0, 0)] * len(key)
                sm_off[c], sm_len[c] = seen[key]
        acc._launch(raw_volume, flip_axes=[], plane=None, k=0, src=list(selection), shift=[(0, 0, 0)] * len(selection),
                    acts=[codes[r] for r in selection], scales=[scales[r] for r in selection], sm_off=sm_off, sm_len=sm_len,
                    sm_src=sm_src, sm_shift=sm_shift, vmask=vmask)

    @staticmethod
    def _validate_patch_first_local_supported(combos, *, image_size, roi_size) -> None:
        """``tta.py:1316-1339``: odd quarter turns need equal image and ROI sizes on the rotated axes."""
        for _f, plane, k in combos:
            if plane is None or k % 2 == 0:
                continue
            axes = tuple(int(a) for a in plane)
            if len({int(image_size[a]) for a in axes}) != 1 or len({int(roi_size[a]) for a in axes}) != 1:
                raise ValueError("Patch-first local TTA only supports odd 90-degree rotations when the "
                                 "rotated axes have equal image and ROI sizes. "
                                 f"Got rotation_plane={tuple(a + 2 for a in axes)}, image_size={tuple(image_size)}, "
                                 f"roi_size={tuple(roi_size)}. Use flip-only TTA, constrain rotations to equal-sized "
                                 "axes such as square XY inputs, or disable "
                                 "`inference.test_time_augmentation.patch_first_local`.")


def __getattr__(name):          # ``from ...inference.tta import TTAPredictor`` as in the reference, without an import cycle
    if name == "TTAPredictor":
        from .tta_predictor import TTAPredictor
        return TTAPredictor
    raise AttributeError(name)


__all__ = ["TTAPredictor", "TTAEnsemble", "TTAEnsembleAccumulator", "apply_view", "resolve_activation_codes", "resolve_activation_specs", "resolve_channel_indices", "resolve_channel_range",
           "resolve_tta_augmentation_combinations", "_resolve_ensemble_mode_map"]
