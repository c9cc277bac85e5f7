"""Affinity-aware inversion of test-time-augmentation views.

Same names and semantics as ``connectomics/inference/tta_affinity.py``: an affinity channel encodes "voxel p and voxel p + offset
belong together", so a flipped / rotated view does not only move voxels — it turns the offset, which may land on ANOTHER
channel of the group (``transform_offset``, ``:72-98``) or on the sign-reversed offset of a channel, in which case the map has to
be re-anchored by a roll whose wrapped face is invalid (``:301-311``, ``:387-391``).  The plan (``build_affinity_tta_plan``,
``:230-329``) is host logic; the inversion itself (``invert_view``, ``:350-393``) is ONE gather kernel here
(``pcb_tta_unview``: un-rotate, un-flip, channel move and roll shift are a single index map) instead of
rot90 + flip + clone + one slice copy per channel.

The reference resolves the affinity channel groups from ``cfg.data.label_transform.targets``
(``data/processing/affinity.py:205-256``); the same walk is restated in ``resolve_affinity_channel_groups_from_cfg`` so that a
tutorial config drives this module unchanged, and explicit ``groups`` / ``mode`` arguments serve callers without a config.
"""

from __future__ import annotations

import ctypes
from collections.abc import Mapping, Sequence
from dataclasses import dataclass
from typing import Any, List, Optional, Tuple, Union

import torch

from .. import _lib as L

ValidityBox = Tuple[slice, ...]
ValidityEntry = Optional[Union[ValidityBox, torch.Tensor]]
AFFINITY_MODES = ("deepem", "banis")


@dataclass(frozen=True)
class ChannelMove:
    """Raw output channel ``src`` becomes canonical channel ``dst``; ``shift`` re-anchors a sign-reversed offset."""

    src: int
    dst: int
    shift: Optional[Tuple[int, ...]] = None


@dataclass(frozen=True)
class AffinityViewPlan:
    moves: Tuple[ChannelMove, ...]
    partial_channels: frozenset

    def shift_for_channel(self, channel: int) -> Optional[Tuple[int, ...]]:
        for mv in self.moves:
            if mv.dst == channel:
                return mv.shift
        return None

    def source_for_channel(self, channel: int) -> int:
        for mv in self.moves:
            if mv.dst == channel:
                return mv.src
        return channel


@dataclass(frozen=True)
class AffinityTTAPlan:
    views: Tuple[AffinityViewPlan, ...]
    partial_channels: frozenset
    shifts: frozenset
    num_channels: int
    spatial_rank: int


@dataclass(frozen=True)
class ViewValidity:
    """Per-channel validity of one canonicalised view: ``None`` (everywhere), a box of slices, or a bool tensor."""

    channels: Tuple[ValidityEntry, ...]

    @classmethod
    def all_valid(cls, num_channels: int) -> "ViewValidity":
        return cls((None,) * int(num_channels))

    def select(self, indices: Optional[Sequence[int]]) -> "ViewValidity":
        if indices is None:
            return self
        return ViewValidity(tuple(self.channels[int(i)] for i in indices))


# ----------------------------------------------------------------------------- config walk (affinity.py:86-256)
def _get(obj: Any, key: str, default: Any = None) -> Any:
    if isinstance(obj, Mapping):
        return obj.get(key, default)
    if hasattr(obj, "get") and not isinstance(obj, (str, bytes)):
        try:
            return obj.get(key, default)
        except TypeError:
            pass
    return getattr(obj, key, default)


def normalize_affinity_mode(mode: Any) -> str:
    if mode is None:
        raise ValueError("Affinity targets require kwargs.affinity_mode: 'deepem' or 'banis'.")
    norm = str(mode).strip().lower()
    if norm not in AFFINITY_MODES:
        raise ValueError(f"Unsupported affinity_mode {mode!r}. Expected one of: {', '.join(AFFINITY_MODES)}.")
    return norm


def parse_affinity_offsets(offsets: Sequence[Any]) -> List[Tuple[int, int, int]]:
    out: List[Tuple[int, int, int]] = []
    for off in offsets:
        if isinstance(off, str):
            parts = off.split("-")
            if len(parts) != 3:
                raise ValueError(f"Invalid affinity offset {off!r}. Expected 'z-y-x' format.")
            out.append((int(parts[0]), int(parts[1]), int(parts[2])))
        elif isinstance(off, (list, tuple)) and len(off) == 3:
            out.append((int(off[0]), int(off[1]), int(off[2])))
        else:
            raise ValueError(f"Unsupported affinity offset {off!r}. Expected 'z-y-x' string or length-3 sequence.")
    return out


def resolve_affinity_offsets_from_kwargs(kwargs: Mapping) -> List[Tuple[int, int, int]]:
    lr = kwargs.get("long_range", None)
    if lr is not None:
        lr = int(lr)
        return [(1, 0, 0), (0, 1, 0), (0, 0, 1), (lr, 0, 0), (0, lr, 0), (0, 0, lr)]
    offsets = kwargs.get("offsets", None)
    if offsets is None or len(offsets) == 0:
        offsets = ["1-0-0", "0-1-0", "0-0-1"]
    return parse_affinity_offsets(offsets)


def _task_name(task: Any) -> Optional[str]:
    if isinstance(task, str):
        return task
    for key in ("name", "task", "type"):
        v = _get(task, key, None)
        if v is not None:
            return v
    return None


def _task_kwargs(task: Any) -> dict:
    raw = _get(task, "kwargs", None) if not isinstance(task, str) else None
    if raw is None:
        return {}
    return dict(raw.items()) if hasattr(raw, "items") else dict(raw)


def _tasks(cfg: Any) -> Tuple[list, bool]:
    label = getattr(getattr(cfg, "data", None), "label_transform", None)
    if label is None:
        return [], True
    targets = getattr(label, "targets", None)
    stack = bool(getattr(label, "stack_outputs", True))
    if targets is None:
        return [], stack
    if isinstance(targets, str):
        return [targets], stack
    try:
        return list(targets), stack
    except TypeError:
        return [], stack


def _stacked_layout(cfg: Any):
    tasks, stack = _tasks(cfg)
    groups: list = []
    if not tasks or not stack:
        return 0, groups
    start = 0
    for task in tasks:
        name, kw = _task_name(task), _task_kwargs(task)
        if name == "affinity":
            offs = resolve_affinity_offsets_from_kwargs(kw)
            width = len(offs)
            groups.append(((start, start + width), offs))
        elif name == "polarity":
            width = 1 if bool(kw.get("exclusive", False)) else 3
        else:
            width = 1
        start += width
    return start, groups


def resolve_affinity_channel_groups_from_cfg(cfg: Any):
    return _stacked_layout(cfg)[1]


def resolve_stacked_label_channel_count(cfg: Any) -> int:
    return _stacked_layout(cfg)[0]


def resolve_affinity_mode_from_cfg(cfg: Any) -> Optional[str]:
    modes = [normalize_affinity_mode(_task_kwargs(t).get("affinity_mode")) for t in _tasks(cfg)[0] if _task_name(t) == "affinity"]
    if not modes:
        return None
    uniq = sorted(set(modes))
    if len(uniq) != 1:
        raise ValueError(f"Mixed affinity_mode values are not supported in one label stack: {uniq}")
    return uniq[0]


# ----------------------------------------------------------------------------- geometry (tta_affinity.py:72-131)
def transform_offset(offset: Sequence[int], *, flip_axes: Sequence[int], rotation_plane_spatial: Optional[Tuple[int, int]],
                     k: int) -> Tuple[int, ...]:
    """Linear part of the inverse rotation followed by the inverse flips, applied to an offset vector."""
    v = [int(x) for x in offset]
    if rotation_plane_spatial is not None:
        p, q = (int(a) for a in rotation_plane_spatial)
        if p == q or min(p, q) < 0 or max(p, q) >= len(v):
            raise ValueError(f"Rotation plane {rotation_plane_spatial} is invalid for an offset with rank {len(v)}.")
        for _ in range((-int(k)) % 4):
            v[p], v[q] = -v[q], v[p]
    for raw in flip_axes:
        ax = int(raw)
        if ax < 0 or ax >= len(v):
            raise ValueError(f"Flip axis {ax} is invalid for an offset with rank {len(v)}.")
        v[ax] = -v[ax]
    return tuple(v)


def valid_slices_for_shift(spatial_shape: Sequence[int], shift: Sequence[int]) -> ValidityBox:
    """Output box holding non-wrapped values after a roll by ``shift``."""
    if len(spatial_shape) != len(shift):
        raise ValueError(f"Roll shift rank {len(shift)} does not match spatial rank {len(spatial_shape)}.")
    box = []
    for size, s in zip(spatial_shape, shift):
        size, s = int(size), int(s)
        if s > 0:
            box.append(slice(min(s, size), size))
        elif s < 0:
            box.append(slice(0, max(0, size + s)))
        else:
            box.append(slice(0, size))
    return tuple(box)


# ----------------------------------------------------------------------------- plan (tta_affinity.py:140-329)
def _resolve_raw_affinity_groups(cfg: Any, *, num_raw: int, requested_head: Optional[str]):
    label_groups = resolve_affinity_channel_groups_from_cfg(cfg)
    if not label_groups:
        return []
    total = resolve_stacked_label_channel_count(cfg)
    model_cfg = getattr(cfg, "model", None)
    heads = _get(model_cfg, "heads", {}) or {}
    if not isinstance(heads, Mapping):
        heads = {}
    if not heads:
        declared = _get(model_cfg, "out_channels", None)
        if declared is None or int(declared) != num_raw or total != num_raw:
            raise ValueError("Affinity TTA requires an unambiguous raw-output to stacked-label mapping. "
                             f"Got model.out_channels={declared}, raw output channels={num_raw}, and "
                             f"stacked label channels={total}; all three must match.")
        window = (0, num_raw)
    else:
        name = requested_head
        if name is None and len(heads) == 1:
            name = next(iter(heads))
        if name is None or name not in heads:
            raise ValueError("Affinity TTA cannot map a named raw output to label channels. Select one "
                             "model head and declare model.heads.<name>.target_slice.")
        head = heads[name]
        tslice = _get(head, "target_slice", None)
        if tslice is not None:
            from .tta import resolve_channel_range
            a, b = resolve_channel_range(tslice, num_channels=total, context=f"model.heads.{name}.target_slice")
            if b - a != num_raw:
                raise ValueError(f"model.heads.{name}.target_slice resolves to width {b - a}, "
                                 f"but the raw output has {num_raw} channels.")
            window = (a, b)
        else:
            width = int(_get(head, "out_channels", 0))
            if len(heads) != 1 or width != num_raw or total != num_raw:
                raise ValueError(f"Affinity TTA cannot prove the label mapping for model head {name!r}. "
                                 f"Got {len(heads)} configured head(s), head out_channels={width}, "
                                 f"raw output channels={num_raw}, and stacked label channels={total}. "
                                 f"Declare model.heads.{name}.target_slice.")
            window = (0, num_raw)
    out = []
    w0, w1 = window
    for (g0, g1), offsets in label_groups:
        lo, hi = max(g0, w0), min(g1, w1)
        if lo >= hi:
            continue
        if len(offsets) != g1 - g0:
            raise ValueError(f"Affinity group [{g0}, {g1}) declares {len(offsets)} offsets; "
                             "its width and offset count must match.")
        out.append(((lo - w0, hi - w0), [tuple(o) for o in offsets[lo - g0:hi - g0]]))
    return out


def build_affinity_tta_plan(cfg: Any = None, *, augmentation_combinations, num_raw: int, requested_head: Optional[str] = None,
                            groups=None, mode: Optional[str] = None) -> Optional[AffinityTTAPlan]:
    """Offset-driven channel moves for every configured view.  ``groups`` (``[((start, stop), [offset, ...]), ...]`` in raw
    output channels) and ``mode`` may be given directly; otherwise they are resolved from ``cfg`` as the reference does."""
    if groups is None:
        if cfg is None or not resolve_affinity_channel_groups_from_cfg(cfg):
            return None
        raw_groups = _resolve_raw_affinity_groups(cfg, num_raw=int(num_raw), requested_head=requested_head)
        mode = resolve_affinity_mode_from_cfg(cfg)
        if mode is None:
            raise ValueError("Affinity channel groups exist but no affinity_mode could be resolved.")
    else:
        raw_groups = [((int(a), int(b)), [tuple(int(v) for v in o) for o in offs]) for (a, b), offs in groups]
        mode = normalize_affinity_mode(mode)
    ranks = {len(o) for _r, offs in raw_groups for o in offs}
    if len(ranks) > 1:
        raise ValueError(f"Mixed affinity offset ranks are not supported: {sorted(ranks)}.")
    rank = next(iter(ranks), 0)
    for rng, offs in raw_groups:
        if len(set(offs)) != len(offs):
            raise ValueError(f"Affinity group {rng} contains duplicate offsets: {offs!r}.")
    views, all_partial, all_shifts = [], set(), set()
    sign = -1 if mode == "banis" else 1
    for flip_axes, plane, k in augmentation_combinations:
        moves, taken = [], set()
        for (a, b), offs in raw_groups:
            if b - a != len(offs):
                raise ValueError(f"Affinity group [{a}, {b}) width does not match its {len(offs)} configured offsets.")
            for si, off in enumerate(offs):
                d = transform_offset(off, flip_axes=flip_axes, rotation_plane_spatial=plane, k=k)
                exact = [i for i, t in enumerate(offs) if t == d]
                rev = [i for i, t in enumerate(offs) if tuple(-x for x in t) == d]
                cand = exact if exact else rev
                if len(cand) != 1:
                    raise ValueError(f"Affinity offset {off} transforms to {d}, but group {offs!r} has {len(cand)} "
                                     f"{'exact' if exact else 'sign-reversed'} counterpart(s).")
                dst = a + cand[0]
                if dst in taken:
                    raise ValueError("Affinity TTA channel mapping is not bijective: multiple source "
                                     f"channels target raw channel {dst}.")
                taken.add(dst)
                shift = None
                if not exact:
                    shift = tuple(sign * int(x) for x in offs[cand[0]])
                    if any(shift):
                        all_partial.add(dst)
                        all_shifts.add(shift)
                    else:
                        shift = None
                moves.append(ChannelMove(src=a + si, dst=dst, shift=shift))
            if {m.dst for m in moves if a <= m.dst < b} != set(range(a, b)):
                raise ValueError(f"Affinity TTA mapping for group [{a}, {b}) is not bijective.")
        views.append(AffinityViewPlan(tuple(moves), frozenset(m.dst for m in moves if m.shift is not None)))
    return AffinityTTAPlan(tuple(views), frozenset(all_partial), frozenset(all_shifts), int(num_raw), rank)


def validate_affinity_output(plan: Optional[AffinityTTAPlan], prediction: torch.Tensor) -> None:
    if plan is None:
        return
    if int(prediction.shape[1]) != plan.num_channels:
        raise ValueError(f"Affinity TTA plan expects {plan.num_channels} raw output channels, "
                         f"but the model produced {int(prediction.shape[1])}.")
    if plan.spatial_rank and plan.spatial_rank != prediction.ndim - 2:
        raise ValueError(f"Affinity offset rank {plan.spatial_rank} does not match raw output spatial "
                         f"rank {prediction.ndim - 2}.")


def view_channel_maps(view_plan: Optional[AffinityViewPlan], num_channels: int):
    """(source channel, roll shift) of every canonical channel for one view — the arrays the kernels take."""
    src = list(range(num_channels))
    shift = [(0, 0, 0)] * num_channels
    if view_plan is not None:
        for mv in view_plan.moves:
            src[mv.dst] = mv.src
            if mv.shift is not None:
                if len(mv.shift) != 3:
                    raise ValueError(f"Affinity roll shift rank {len(mv.shift)} does not match raw output spatial rank 3.")
                shift[mv.dst] = tuple(int(v) for v in mv.shift)
    return src, shift


def invert_view(prediction: torch.Tensor, *, flip_axes: Sequence[int], rotation_plane_spatial: Optional[Tuple[int, int]], k: int,
                view_plan: Optional[AffinityViewPlan], tta_plan: Optional[AffinityTTAPlan]):
    """``(canonical prediction, ViewValidity)`` of one view: rot90(-k), flip, channel moves and roll shifts in one gather."""
    L.require_device(prediction, "TTA invert_view")
    if prediction.dim() != 5:
        raise ValueError(f"pcb200 TTA expects [N,C,D,H,W] predictions; got shape {tuple(prediction.shape)}")
    pred = prediction.contiguous()
    n, c = int(pred.shape[0]), int(pred.shape[1])
    vsize = [int(v) for v in pred.shape[2:]]
    kk = int(k) % 4 if rotation_plane_spatial is not None else 0
    ra, rb = (-1, -1) if rotation_plane_spatial is None else (int(rotation_plane_spatial[0]), int(rotation_plane_spatial[1]))
    size = list(vsize)
    if ra >= 0 and (kk & 1):
        size[ra], size[rb] = vsize[rb], vsize[ra]
    out = torch.empty((n, c, *size), device=pred.device, dtype=pred.dtype)
    validate_affinity_output(tta_plan, out)
    src, shift = view_channel_maps(view_plan, c)
    fm = 0
    for a in (flip_axes or []):
        fm |= 1 << int(a)
    flat = [v for s in shift for v in s]
    with torch.cuda.device(pred.device):
        L.check(L.lib().pcb_tta_unview(L.ptr(pred), L.ptr(out), L.dtype_code(pred.dtype), ctypes.c_int64(n), ctypes.c_int64(c),
                                       ctypes.c_int64(c), L.i64x(size), fm, ra, rb, kk, (ctypes.c_int * c)(*src),
                                       (ctypes.c_int * (3 * c))(*flat), L.stream_ptr(pred.device)), "pcb_tta_unview")
    validity: List[ValidityEntry] = [None] * c
    for ch in range(c):
        if any(shift[ch]):
            validity[ch] = valid_slices_for_shift(size, shift[ch])
    return out, ViewValidity(tuple(validity))


__all__ = ["AffinityTTAPlan", "AffinityViewPlan", "ChannelMove", "ViewValidity", "build_affinity_tta_plan", "invert_view",
           "transform_offset", "valid_slices_for_shift", "validate_affinity_output", "view_channel_maps",
           "resolve_affinity_channel_groups_from_cfg", "resolve_affinity_mode_from_cfg", "resolve_stacked_label_channel_count"]
