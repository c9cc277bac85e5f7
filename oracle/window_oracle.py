"""CPU oracle for the sliding-window path (grid, weight maps, crop/pad, blend, normalise).

TEST INFRASTRUCTURE — never imported by the product package.

Restates, in plain numpy / torch-CPU, the algorithm of the reference's
``connectomics/inference/window.py`` (eager engine) and the integer grid /
accumulate part of ``connectomics/inference/lazy.py``.  Each function cites the lines it
follows.  PINNED: ``oracle/make_goldens.py`` runs the real reference ``window.py``
(file-loaded from /root/reference, possible in the build container only) and stores its
outputs in ``tests/golden/window_goldens.npz``; ``tests/test_oracle_window.py`` checks
this restatement against those vectors and against the known answers in the reference's
own tests (``tests/unit/test_lazy_inference.py:73-84``, ``tests/unit/test_window_engine.py:33-144``).
"""

from __future__ import annotations

import itertools
from typing import Callable, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

DT_MODES = {"distance", "distance_transform", "distance-transform", "distance_transform_cdt",
            "banis", "banis_distance"}  # window.py:30-37


def scan_interval(image_size, roi_size, overlap) -> Tuple[int, ...]:
    """window.py:57-89 — stride per axis; Python round() (banker's) on roi*(1-ov)."""
    nd = len(roi_size)
    ov = [float(overlap[i]) for i in range(nd)] if isinstance(overlap, (list, tuple)) \
        else [float(overlap)] * nd
    out = []
    for a in range(nd):
        o = min(max(ov[a], 0.0), 0.99)
        roi, img = int(roi_size[a]), int(image_size[a])
        out.append(img if img <= roi else max(1, int(round(roi * (1.0 - o)))))
    return tuple(out)


def axis_starts(img: int, roi: int, stride: int) -> List[int]:
    """window.py:107-118 — starts along one axis, last one snapped to img-roi."""
    if img <= roi:
        return [0]
    s = list(range(0, img - roi + 1, max(1, stride)))
    if s[-1] != img - roi:
        s.append(img - roi)
    return s


def dense_starts(image_size, roi_size, interval) -> List[Tuple[int, ...]]:
    """window.py:92-134 — z-major Cartesian product of the per-axis starts."""
    per_axis = [axis_starts(int(image_size[a]), int(roi_size[a]), int(interval[a]))
                for a in range(len(roi_size))]
    return [tuple(p) for p in itertools.product(*per_axis)]


def lazy_axis_offsets(image_size, roi_size, overlap, snap_to_edge: bool) -> List[List[int]]:
    """lazy.py:269-334 — lazy grid with face-centred boundary windows (negative starts)."""
    out = []
    if snap_to_edge:
        strides = [max(1, int(int(roi_size[a]) * (1.0 - float(overlap[a])))) for a in range(3)]
    else:
        strides = list(scan_interval(image_size, roi_size, tuple(float(v) for v in overlap)))
    for a in range(3):
        img, roi, st = int(image_size[a]), int(roi_size[a]), int(strides[a])
        bp = max(0, roi - st)
        if img <= roi:
            out.append([0])
            continue
        st = max(1, st)
        lo, hi = -bp, img - roi + bp
        offs = list(range(lo, hi + 1, st))
        if not offs or offs[-1] != hi:
            offs.append(hi)
        out.append(offs)
    return out


def lazy_region_records(image_size, roi_size, overlap, region_start, region_stop, snap_to_edge):
    """lazy.py:337-365 + 1077-1102 — windows intersecting a region and their clipped boxes.

    Returns list of (patch_start, pred_lo, pred_hi, out_lo, out_hi) integer triples."""
    per_axis = lazy_axis_offsets(image_size, roi_size, overlap, snap_to_edge)
    keep = []
    for a, offs in enumerate(per_axis):
        keep.append([o for o in offs
                     if o < int(region_stop[a]) and o + int(roi_size[a]) > int(region_start[a])])
    recs = []
    for ps in itertools.product(*keep):
        lo = tuple(max(ps[a], int(region_start[a])) for a in range(3))
        hi = tuple(min(ps[a] + int(roi_size[a]), int(region_stop[a])) for a in range(3))
        if any(hi[a] <= lo[a] for a in range(3)):
            continue
        recs.append((tuple(ps),
                     tuple(lo[a] - ps[a] for a in range(3)), tuple(hi[a] - ps[a] for a in range(3)),
                     tuple(lo[a] - int(region_start[a]) for a in range(3)),
                     tuple(hi[a] - int(region_start[a]) for a in range(3))))
    return recs


def importance_map(roi_size, mode: str, dtype=torch.float32, min_value: float = 1e-5) -> torch.Tensor:
    """window.py:137-243 — separable bump / constant / distance-transform map in `dtype`."""
    mode = str(mode).strip().lower()
    roi = tuple(int(v) for v in roi_size)
    if any(v <= 0 for v in roi):
        raise ValueError("roi_size must be positive")
    tiny = torch.finfo(dtype).tiny
    if mode in DT_MODES:
        m = None
        for a, n in enumerate(roi):
            c = torch.arange(n, dtype=dtype)
            d = torch.minimum(c + 1, torch.as_tensor(n, dtype=dtype) - c)
            shape = [1] * len(roi)
            shape[a] = n
            d = d.reshape(shape)
            m = d if m is None else torch.minimum(m, d)
        return m
    if mode == "constant":
        m = torch.ones(roi, dtype=dtype)
    elif mode == "bump":
        m = None
        for a, n in enumerate(roi):
            i = torch.arange(n, dtype=dtype)
            u = (i + 1.0) / (n + 1.0) * 2.0 - 1.0
            k = torch.exp(-1.0 / (1.0 - u * u).clamp_min(tiny))
            k = k / k.max().clamp_min(tiny)
            shape = [1] * len(roi)
            shape[a] = n
            m = k.view(shape) if m is None else m * k.view(shape)
        m = m.clamp_min(tiny)
    else:
        raise ValueError(f"unsupported blending mode {mode!r}")
    return m.clamp_min(min_value) if min_value > 0 else m


def normalize_accumulator(value: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """window.py:275-294 — value /= clamp_min(weight, 1e-4) in the value dtype (in place)."""
    floor = 1.0e-4
    if value.dtype == torch.float16:
        floor = max(floor, float(torch.finfo(torch.float16).tiny))
    div = torch.clamp_min(weight, floor).to(value.dtype)
    value /= div
    return value


def extract_patch(vol: torch.Tensor, start, roi, padding_mode: str, cval: float) -> torch.Tensor:
    """window.py:464-527 — crop [1,C,*roi] at `start` (may stick out), pad the missing part;
    reflect/circular fall back to constant when a pad >= the cropped dim."""
    nd = len(roi)
    img = vol.shape[-nd:]
    end = [start[a] + roi[a] for a in range(nd)]
    lo = [max(0, start[a]) for a in range(nd)]
    hi = [min(int(img[a]), end[a]) for a in range(nd)]
    inner = vol[(slice(None), slice(None)) + tuple(slice(lo[a], hi[a]) for a in range(nd))]
    pads = [(max(0, -start[a]), max(0, end[a] - int(img[a]))) for a in range(nd)]
    if any(b or e for b, e in pads):
        mode = padding_mode
        if mode in ("reflect", "circular"):
            dims = inner.shape[-nd:]
            if any(pads[a][0] >= int(dims[a]) or pads[a][1] >= int(dims[a]) for a in range(nd)):
                mode = "constant"
        flat = []
        for b, e in reversed(pads):
            flat += [b, e]
        inner = F.pad(inner, tuple(flat), mode=mode, value=cval) if mode == "constant" \
            else F.pad(inner, tuple(flat), mode=mode)
    return inner


def eager_sliding_window(inputs: torch.Tensor, network: Callable, roi_size, overlap=0.5,
                         mode="bump", padding_mode="constant", cval=0.0, sw_batch_size=1):
    """window.py:563-683 — the eager engine: grow-to-ROI, probe window first, batches of
    sw_batch_size, value += out*w, weight += w, normalise, crop."""
    roi = tuple(int(v) for v in roi_size)
    nd = len(roi)
    if inputs.dim() < nd + 2:
        raise ValueError("inputs must be (B, C, *spatial)")
    if inputs.shape[0] != 1:
        raise ValueError("batch size must be 1")
    orig = tuple(int(v) for v in inputs.shape[-nd:])
    grow = [max(0, roi[a] - orig[a]) for a in range(nd)]
    if any(grow):
        flat = []
        for g in reversed(grow):
            flat += [0, g]
        inputs = F.pad(inputs, tuple(flat), mode="constant", value=cval)
    img = tuple(int(v) for v in inputs.shape[-nd:])
    starts = dense_starts(img, roi, scan_interval(img, roi, overlap))
    probe = network(extract_patch(inputs, starts[0], roi, padding_mode, cval))
    if not isinstance(probe, torch.Tensor):
        raise ValueError("network must return a tensor")
    cout, dt = int(probe.shape[1]), probe.dtype
    w = importance_map(roi, mode, dtype=dt)
    wb = w.view(1, 1, *roi)
    val = torch.zeros((1, cout, *img), dtype=dt)
    wacc = torch.zeros((1, 1, *img), dtype=dt)

    def acc(out1, st):
        idx = (slice(None), slice(None)) + tuple(slice(st[a], st[a] + roi[a]) for a in range(nd))
        val[idx] += out1.to(dt) * wb
        wacc[idx] += wb

    acc(probe[0:1], starts[0])
    rest = starts[1:]
    for b0 in range(0, len(rest), sw_batch_size):
        chunk = rest[b0:b0 + sw_batch_size]
        batch = torch.cat([extract_patch(inputs, s, roi, padding_mode, cval) for s in chunk], 0)
        out = network(batch)
        for i, s in enumerate(chunk):
            acc(out[i:i + 1], s)
    res = normalize_accumulator(val, wacc)
    if any(grow):
        res = res[(slice(None), slice(None)) + tuple(slice(0, orig[a]) for a in range(nd))].contiguous()
    return res


def lazy_sliding_window(volume: torch.Tensor, network: Callable, roi_size, overlap, mode="bump",
                        padding_mode="constant", cval=0.0, region_start=None, region_stop=None,
                        snap_to_edge=False, out_dtype=torch.float32, sw_batch_size=1,
                        rank=0, world_size=1, normalize=True):
    """lazy.py:986-1258 on an in-memory volume [1,C,D,H,W]: lazy grid ∩ region, optional
    [rank::world] shard, fp32 patches -> network -> out_dtype, accumulate clipped boxes."""
    roi = tuple(int(v) for v in roi_size)
    ov = tuple(float(v) for v in overlap) if isinstance(overlap, (list, tuple)) else (float(overlap),) * 3
    img = tuple(int(v) for v in volume.shape[-3:])
    start = (0, 0, 0) if region_start is None else tuple(max(0, int(v)) for v in region_start)
    stop = img if region_stop is None else tuple(min(img[a], int(region_stop[a])) for a in range(3))
    osz = tuple(stop[a] - start[a] for a in range(3))
    recs = lazy_region_records(img, roi, ov, start, stop, snap_to_edge)[rank::world_size]
    w = importance_map(roi, mode, dtype=out_dtype).view(1, 1, *roi)
    val = None
    wacc = torch.zeros((1, 1, *osz), dtype=out_dtype)
    for b0 in range(0, len(recs), sw_batch_size):
        chunk = recs[b0:b0 + sw_batch_size]
        batch = torch.cat([extract_patch(volume, r[0], roi, padding_mode, cval) for r in chunk], 0)
        pred = network(batch.float()).detach().to(out_dtype)
        if val is None:
            val = torch.zeros((1, int(pred.shape[1]), *osz), dtype=out_dtype)
        for i, (_, plo, phi, olo, ohi) in enumerate(chunk):
            ps = tuple(slice(plo[a], phi[a]) for a in range(3))
            os_ = tuple(slice(olo[a], ohi[a]) for a in range(3))
            val[(slice(None), slice(None)) + os_] += pred[(slice(i, i + 1), slice(None)) + ps] * \
                w[(slice(None), slice(None)) + ps]
            wacc[(slice(None), slice(None)) + os_] += w[(slice(None), slice(None)) + ps]
    if not normalize:
        return val, wacc
    return normalize_accumulator(val, wacc)


def chunk_grid(volume_shape, chunk_shape):
    """chunked/chunk_grid.py:32-43 — ceil-div grid of (index, key, start, stop)."""
    counts = [-(-int(volume_shape[a]) // int(chunk_shape[a])) for a in range(3)]
    out = []
    for idx in itertools.product(*[range(c) for c in counts]):
        st = tuple(idx[a] * int(chunk_shape[a]) for a in range(3))
        sp = tuple(min(int(volume_shape[a]), st[a] + int(chunk_shape[a])) for a in range(3))
        out.append((idx, f"z{idx[0]}_y{idx[1]}_x{idx[2]}", st, sp))
    return out


def binary_jaccard(pred: np.ndarray, target: np.ndarray, threshold: float = 0.5) -> float:
    """evaluation/metric_execution.py:178-196 semantics: IoU of (pred > thr) vs (target > 0)."""
    p = np.asarray(pred) > threshold
    t = np.asarray(target) > 0
    union = np.logical_or(p, t).sum()
    return float(np.logical_and(p, t).sum() / union) if union else 1.0
