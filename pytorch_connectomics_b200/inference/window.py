"""Sliding-window tiled inference on the B200 engine — drop-in for
``connectomics/inference/window.py`` (same public names, argument meaning and errors).

Integer grid logic runs in the C library on the host (``pcb_sw_scan_interval`` / ``pcb_sw_plan``,
bit-exact with ``window.py:57-134``); the weight map, crop+pad, overlap-add and normalise steps are
CUDA kernels behind ``include/pcb200.h``.  Windows are accumulated one launch per window in grid
order on one stream, so every voxel's fp sum associates exactly like the reference's sequential
``value[loc] += out * w`` — for the same ``network`` outputs the blended volume is bit-identical.

Tensor-producing functions need CUDA tensors (no CPU fallback); pure-integer helpers work anywhere.
"""

from __future__ import annotations

import ctypes
import logging
from collections.abc import Mapping
from typing import Callable, List, Optional, Sequence, Tuple, Union

import torch

from .. import _lib as L

logger = logging.getLogger(__name__)

_DISTANCE_TRANSFORM_BLEND_MODES = {"distance", "distance_transform", "distance-transform",
                                   "distance_transform_cdt", "banis", "banis_distance"}


def _cfg_value(obj, key: str, default=None):
    if obj is None:
        return default
    return obj.get(key, default) if isinstance(obj, Mapping) else getattr(obj, key, default)


def _normalize_blending_mode(mode: str) -> str:
    return str(mode).strip().lower()


def is_distance_transform_blending(mode: str) -> bool:
    return _normalize_blending_mode(mode) in _DISTANCE_TRANSFORM_BLEND_MODES


def _overlap3(overlap, nd: int) -> List[float]:
    if isinstance(overlap, (list, tuple)):
        return [float(overlap[i]) for i in range(nd)]
    return [float(overlap)] * nd


def _pad3(vals: Sequence[int], fill: int) -> List[int]:
    """Left-pad a 1-/2-/3-D size to 3-D (the C ABI is 3-D; lower ranks get leading singleton axes)."""
    v = [int(x) for x in vals]
    if len(v) > 3:
        raise ValueError(f"at most 3 spatial dims are supported, got {len(v)}")
    return [fill] * (3 - len(v)) + v


# ----------------------------------------------------------------------------- integer grid (host)
def compute_scan_interval(image_size, roi_size, num_spatial_dims: Optional[int] = None,
                          overlap: Union[float, Sequence[float]] = 0.0) -> Tuple[int, ...]:
    """``window.py:57-89``."""
    del num_spatial_dims
    nd = len(roi_size)
    ov = _overlap3(overlap, nd)
    out = (ctypes.c_int64 * 3)()
    L.check(L.lib().pcb_sw_scan_interval(L.i64x(_pad3(image_size[:nd], 1)), L.i64x(_pad3(roi_size, 1)),
                                         L.f64x([0.0] * (3 - nd) + ov), out), "pcb_sw_scan_interval")
    return tuple(int(v) for v in out)[3 - nd:]


def _plan(kind: int, image_size, roi_size, overlap, region=None) -> List[Tuple[int, ...]]:
    nd = len(roi_size)
    img, roi = _pad3(image_size[:nd], 1), _pad3(roi_size, 1)
    ov = L.f64x([0.0] * (3 - nd) + _overlap3(overlap, nd))
    reg = None
    if region is not None:
        lo, hi = region
        reg = L.i64x(_pad3(lo, 0) + _pad3(hi, 1))
    cnt = ctypes.c_int64(0)
    lib = L.lib()
    L.check(lib.pcb_sw_plan(kind, L.i64x(img), L.i64x(roi), ov, reg, None, ctypes.c_int64(0), ctypes.byref(cnt)),
            "pcb_sw_plan")
    n = int(cnt.value)
    buf = (ctypes.c_int64 * (3 * max(n, 1)))()
    L.check(lib.pcb_sw_plan(kind, L.i64x(img), L.i64x(roi), ov, reg, buf, ctypes.c_int64(n), ctypes.byref(cnt)),
            "pcb_sw_plan")
    return [tuple(int(buf[3 * i + a]) for a in range(3 - nd, 3)) for i in range(n)]


def dense_patch_slices(image_size, roi_size, scan_interval, return_slice: bool = True):
    """``window.py:92-134`` — window starts in z-major order, last start snapped to ``img - roi``.
    ``scan_interval`` is honoured by re-deriving the grid from it (overlap = 1 - stride/roi exact
    only for the engine's own intervals, so the starts are built here from the given strides)."""
    nd = len(roi_size)
    per_axis: List[List[int]] = []
    for a in range(nd):
        roi, img, st = int(roi_size[a]), int(image_size[a]), max(1, int(scan_interval[a]))
        if img <= roi:
            per_axis.append([0])
            continue
        s = list(range(0, img - roi + 1, st))
        if s[-1] != img - roi:
            s.append(img - roi)
        per_axis.append(s)
    starts: List[Tuple[int, ...]] = [()]
    for axis_starts in per_axis:
        starts = [p + (s,) for p in starts for s in axis_starts]
    if not return_slice:
        return starts
    return [tuple(slice(s, s + int(roi_size[i])) for i, s in enumerate(st)) for st in starts]


# ----------------------------------------------------------------------------- weight maps (device)
def _device_or_raise(device) -> torch.device:
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"pcb200: sliding-window kernels need a CUDA device (sm_100a); got {dev}. "
                           "There is no CPU fallback for this path.")
    return dev


def _imap(roi_size, blend: int, device, dtype, min_value: float) -> torch.Tensor:
    roi = tuple(int(v) for v in roi_size)
    if not roi or any(v <= 0 for v in roi):
        raise ValueError(f"roi_size must contain positive values, got {roi_size}.")
    dev = _device_or_raise(device)
    out = torch.empty(roi, device=dev, dtype=dtype)
    with torch.cuda.device(dev):
        L.check(L.lib().pcb_sw_importance_map(blend, L.i64x(_pad3(roi, 1)), len(roi), L.dtype_code(dtype),
                                              ctypes.c_double(min_value), L.ptr(out), L.stream_ptr(dev)),
                "pcb_sw_importance_map")
    return out


def compute_importance_map(roi_size, *, mode: str = "constant", device="cuda",
                           dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """``window.py:137-196`` — ``constant`` or Wu ``bump`` map, floor-clamped to ``finfo.tiny``."""
    m = _normalize_blending_mode(mode)
    if m not in ("constant", "bump"):
        if any(int(v) <= 0 for v in roi_size):
            raise ValueError(f"roi_size must contain positive values, got {roi_size}.")
        raise ValueError(f"compute_importance_map: unsupported mode {mode!r}; expected 'constant' or 'bump' "
                         "(use is_distance_transform_blending for the distance-transform path).")
    return _imap(roi_size, L.BLEND[m], device, dtype, 0.0)


def build_sliding_importance_map(roi_size, *, mode: str, device, dtype: torch.dtype = torch.float32,
                                 min_value: float = 1e-5) -> torch.Tensor:
    """``window.py:199-243`` — bump/constant maps floored at ``min_value``; distance-transform map."""
    m = _normalize_blending_mode(mode)
    if m in _DISTANCE_TRANSFORM_BLEND_MODES:
        return _imap(roi_size, L.BLEND["distance"], device, dtype, 0.0)
    if m not in ("constant", "bump"):
        return compute_importance_map(roi_size, mode=m, device=device, dtype=dtype)  # raises
    return _imap(roi_size, L.BLEND[m], device, dtype, float(min_value))


def build_sliding_accumulator_weight_maps(roi_size, *, mode: str, device, value_dtype: torch.dtype):
    """``window.py:246-272`` — value map and weight map are the SAME tensor (same dtype)."""
    m = build_sliding_importance_map(roi_size, mode=mode, device=device, dtype=value_dtype)
    return m, m


def normalize_weighted_accumulator(value_accumulator: torch.Tensor, weight_accumulator: torch.Tensor) -> torch.Tensor:
    """``window.py:275-294`` — in place ``value /= clamp_min(weight, 1e-4)`` in the value dtype."""
    L.require_device(value_accumulator, "normalize_weighted_accumulator")
    if not (value_accumulator.is_contiguous() and weight_accumulator.is_contiguous()):
        raise ValueError("normalize_weighted_accumulator expects contiguous accumulators")
    if weight_accumulator.dtype != value_accumulator.dtype:
        weight_accumulator = weight_accumulator.to(value_accumulator.dtype)
    nvox = weight_accumulator.numel()
    if nvox == 0 or value_accumulator.numel() % nvox:
        raise ValueError("accumulator shapes do not broadcast")
    cout = value_accumulator.numel() // nvox
    with torch.cuda.device(value_accumulator.device):
        L.check(L.lib().pcb_sw_normalize(L.ptr(value_accumulator), L.ptr(weight_accumulator),
                                         L.dtype_code(value_accumulator.dtype), ctypes.c_int64(cout),
                                         ctypes.c_int64(nvox), L.stream_ptr(value_accumulator.device)),
                "pcb_sw_normalize")
    return value_accumulator


def apply_border_mask(importance_map: torch.Tensor, border_mask: Sequence[int]) -> torch.Tensor:
    """``window.py:297-319`` — zero the outer k voxels per axis (lazy path only)."""
    if not border_mask or all(int(b) <= 0 for b in border_mask):
        return importance_map
    nd = len(border_mask)
    shape = importance_map.shape[-nd:]
    for axis, k in enumerate(int(b) for b in border_mask):
        if k <= 0:
            continue
        size = int(shape[axis])
        if 2 * k >= size:
            raise ValueError(f"inference.sliding_window.border_mask[{axis}]={k} is too large "
                             f"for window size {size} on that axis.")
        dim = importance_map.ndim - nd + axis
        importance_map.narrow(dim, 0, k).zero_()
        importance_map.narrow(dim, size - k, k).zero_()
    return importance_map


# ----------------------------------------------------------------------------- config resolvers
_DTYPE_ALIASES = {"float32": torch.float32, "fp32": torch.float32, "float16": torch.float16, "fp16": torch.float16,
                  "half": torch.float16, "bfloat16": torch.bfloat16, "bf16": torch.bfloat16}


def resolve_model_output_dtype(cfg) -> torch.dtype:
    """``window.py:333-346``."""
    raw = _cfg_value(_cfg_value(_cfg_value(cfg, "inference"), "model"), "output_dtype")
    if raw is None:
        return torch.float32
    name = str(raw).strip().lower().removeprefix("torch.")
    if name not in _DTYPE_ALIASES:
        raise ValueError(f"inference.model.output_dtype must be one of {sorted(_DTYPE_ALIASES)}, got {raw!r}.")
    return _DTYPE_ALIASES[name]


def _sliding_cfg(cfg):
    return getattr(getattr(cfg, "inference", None), "sliding_window", None)


def resolve_border_mask(cfg, spatial_dims: int) -> List[int]:
    """``window.py:349-363``."""
    raw = getattr(_sliding_cfg(cfg), "border_mask", None)
    if not raw:
        return []
    vals = [int(v) for v in raw]
    if len(vals) == 1:
        vals = vals * spatial_dims
    if len(vals) != spatial_dims:
        raise ValueError(f"inference.sliding_window.border_mask must have length 1 or {spatial_dims}, got {len(vals)}.")
    return vals


def is_2d_inference_mode(cfg) -> bool:
    data = getattr(cfg, "data", None)
    return bool(getattr(getattr(data, "train", None), "do_2d", False) or getattr(getattr(data, "val", None), "do_2d", False))


def resolve_inferer_roi_size(cfg) -> Optional[Tuple[int, ...]]:
    """``window.py:373-396`` — window_size, else model.output_size, else data patch_size."""
    ws = getattr(_sliding_cfg(cfg), "window_size", None)
    if ws:
        return tuple(int(v) for v in ws)
    for holder, key in ((getattr(cfg, "model", None), "output_size"),
                        (getattr(getattr(cfg, "data", None), "data_transform", None), "patch_size")):
        size = getattr(holder, key, None) if holder is not None else None
        if size:
            roi = tuple(int(v) for v in size)
            if is_2d_inference_mode(cfg) and len(roi) == 2:
                roi = (1,) + roi
            return roi
    return None


def resolve_inferer_overlap(cfg, roi_size) -> Union[float, Tuple[float, ...]]:
    """``window.py:399-410`` — default 0.5, clamped to [0, 0.99]."""
    sc = _sliding_cfg(cfg)
    ov = getattr(sc, "overlap", None) if sc is not None else None
    if ov is None:
        return 0.5
    clamp = lambda o: float(max(0.0, min(o, 0.99)))  # noqa: E731
    return tuple(clamp(o) for o in ov) if isinstance(ov, (list, tuple)) else clamp(ov)


def _none_if_blank(v):
    return None if isinstance(v, str) and v.lower() in {"", "none", "null"} else v


def _resolve_sliding_window_runtime(cfg, roi_size) -> dict:
    """``window.py:413-461``."""
    sc = _sliding_cfg(cfg)
    loader = getattr(getattr(cfg, "data", None), "dataloader", None)
    batch_default = getattr(loader, "batch_size", 1) if loader else 1
    cfg_bs = getattr(sc, "sw_batch_size", None) if sc else None
    rt = {
        "overlap": resolve_inferer_overlap(cfg, roi_size),
        "sw_batch_size": max(1, int(cfg_bs if cfg_bs is not None else batch_default)),
        "mode": _normalize_blending_mode(getattr(sc, "blending", "bump") if sc else "bump"),
        "padding_mode": getattr(sc, "padding_mode", "constant") if sc else "constant",
        "cval": float(getattr(sc, "cval", 0.0)) if sc else 0.0,
        "keep_input_on_cpu": bool(getattr(sc, "keep_input_on_cpu", False)) if sc else False,
        "sw_device": _none_if_blank(getattr(sc, "sw_device", None) if sc else None),
        "output_device": _none_if_blank(getattr(sc, "output_device", None) if sc else None),
    }
    if rt["keep_input_on_cpu"]:
        if rt["sw_device"] is None and torch.cuda.is_available():
            rt["sw_device"] = "cuda"
        if rt["output_device"] is None:
            rt["output_device"] = "cpu"
    return rt


# ----------------------------------------------------------------------------- crop + pad (device)
def _slice_starts(patch_slices) -> List[Tuple[int, ...]]:
    return [tuple(int(s.start) for s in ps) for ps in patch_slices]


def _extract_padded_patch_batch(tensor: torch.Tensor, patch_slices, *, roi_size, padding_mode: str, cval: float):
    """``window.py:464-527`` — [n, C, *roi] windows gathered (and padded) in one kernel launch."""
    if tensor.shape[0] != 1:
        raise ValueError("Patch-first sliding-window TTA currently expects singleton batches. "
                         f"Got batch size {tensor.shape[0]}.")
    locations = _slice_starts(patch_slices)
    return _extract_starts(tensor, locations, tuple(int(v) for v in roi_size), padding_mode, cval), locations


def _extract_starts(tensor: torch.Tensor, starts, roi, padding_mode: str, cval: float) -> torch.Tensor:
    L.require_device(tensor, "sliding-window patch extraction")
    if padding_mode not in L.PAD:
        raise ValueError(f"unsupported padding_mode {padding_mode!r}; expected one of {sorted(L.PAD)}")
    nd = len(roi)
    vol = tensor.contiguous()
    c = int(vol.shape[1])
    img = [int(v) for v in vol.shape[-nd:]]
    n = len(starts)
    out = torch.empty((n, c, *roi), device=vol.device, dtype=vol.dtype)
    flat = []
    for s in starts:
        flat += _pad3(s, 0)
    with torch.cuda.device(vol.device):
        L.check(L.lib().pcb_sw_extract(L.ptr(vol), L.dtype_code(vol.dtype), ctypes.c_int64(c), L.i64x(_pad3(img, 1)),
                                       L.i64x(_pad3(roi, 1)), L.i64x(flat), ctypes.c_int64(n),
                                       ctypes.c_int(L.PAD[padding_mode]), ctypes.c_double(cval), L.ptr(out),
                                       L.stream_ptr(vol.device)), "pcb_sw_extract")
    return out


def _accumulate_window(pred: torch.Tensor, wmap: torch.Tensor, value: torch.Tensor, weight: torch.Tensor,
                       roi, out_size, pred_lo, out_lo, box) -> None:
    """One window's ``value[box] += pred[box]*map; weight[box] += map`` (``window.py:648-655``)."""
    with torch.cuda.device(value.device):
        L.check(L.lib().pcb_sw_accumulate(L.ptr(pred), L.ptr(wmap), L.ptr(value), L.ptr(weight),
                                          L.dtype_code(value.dtype), ctypes.c_int64(int(value.shape[1])),
                                          L.i64x(_pad3(roi, 1)), L.i64x(_pad3(out_size, 1)), L.i64x(_pad3(pred_lo, 0)),
                                          L.i64x(_pad3(out_lo, 0)), L.i64x(_pad3(box, 1)), L.stream_ptr(value.device)),
                "pcb_sw_accumulate")


def _accumulate_batch(pred: torch.Tensor, wmap: torch.Tensor, value: torch.Tensor, weight: torch.Tensor,
                      roi, out_size, starts) -> None:
    """``window.py:641-655`` for one network batch: every full window of ``pred`` [n,Cout,*roi] in ONE launch per 16
    windows, bit-identical to the per-window loop in list order (``pcb_sw_accumulate_batch``)."""
    flat = []
    for s in starts:
        flat += _pad3(s, 0)
    with torch.cuda.device(value.device):
        L.check(L.lib().pcb_sw_accumulate_batch(L.ptr(pred), L.ptr(wmap), L.ptr(value), L.ptr(weight),
                                                L.dtype_code(value.dtype), ctypes.c_int64(int(value.shape[1])),
                                                L.i64x(_pad3(roi, 1)), L.i64x(_pad3(out_size, 1)), L.i64x(flat),
                                                ctypes.c_int64(len(starts)), L.stream_ptr(value.device)),
                "pcb_sw_accumulate_batch")


def check_network_output(out, n: int, cout: Optional[int], roi, who: str) -> torch.Tensor:
    """The kernels read ``out`` through raw pointers as [n, cout, *roi]; the reference would fail with a broadcast error
    for anything else (``window.py:648-655``), so refuse it here instead of reading out of bounds."""
    if not isinstance(out, torch.Tensor):
        raise ValueError(f"{who}: `network` must return a torch.Tensor; got {type(out).__name__}.")
    want_tail = tuple(int(v) for v in roi)
    ok = out.dim() == len(want_tail) + 2 and int(out.shape[0]) == n and tuple(int(v) for v in out.shape[2:]) == want_tail
    if ok and cout is not None:
        ok = int(out.shape[1]) == cout
    if not ok:
        want = (n, cout if cout is not None else "Cout") + want_tail
        raise ValueError(f"{who}: `network` returned shape {tuple(out.shape)} for a batch of {n} windows; expected "
                         f"{want} (window-sized outputs with a fixed channel count).")
    return out


def native_plan_of(network):
    """The ``pcb_net`` plan behind ``network`` when it is a pcb200 MedNeXt (module or wrapper) in inference mode, else
    ``None``.  Arbitrary callables (lambdas, TTA closures, other frameworks' modules) run through the generic loop."""
    getter = getattr(network, "native_plan", None)
    if getter is None or not callable(getter):
        return None
    try:
        return getter()
    except Exception:          # building the plan is an optimisation, never a reason to fail the inference call
        logger.debug("native plan unavailable", exc_info=True)
        return None


def run_window_list(vol: torch.Tensor, network, starts, *, roi, image, sw_batch_size: int, padding_mode: str, cval: float,
                    mode: str, sw_device, work_device, value=None, weight=None, wmap=None, probe_first: bool = True,
                    cuda_graph: bool = True, who: str = "sliding window"):
    """The tile loop of ``window.py:603-655`` over ``starts`` (grid order) on a device-resident ``vol`` [1,C,*image]:
    crop+pad -> network -> ``value += pred*map; weight += map`` into the given (or new, zeroed) accumulators.

    * ``network`` is a pcb200 MedNeXt: ONE library call (``pcb_sw_run``) runs every batch — crop, the whole network and
      the blend are enqueued natively, as a replayed CUDA graph per batch when ``cuda_graph``;
    * any other callable: crop kernel -> ``network(batch)`` -> one blend launch per batch.
    Both blend in list order with the same arithmetic, so the accumulators are bit-identical for identical predictions."""
    roi = tuple(int(v) for v in roi)
    state = {"value": value, "weight": weight, "wmap": wmap}
    plan = native_plan_of(network) if (len(roi) == 3 and vol.is_cuda and vol.device == torch.device(work_device)
                                       and torch.device(sw_device) == vol.device) else None
    if plan is not None and starts and int(vol.shape[1]) == plan.in_channels and all(r % 16 == 0 for r in roi) \
            and padding_mode in L.PAD and len(plan.head_channels) == 1 and plan.device == vol.device:
        odt = plan.model._odt(vol)
        if state["value"] is None:
            state["wmap"] = build_sliding_importance_map(roi, mode=mode, device=work_device, dtype=odt)
            state["value"] = torch.zeros((1, plan.head_channels[0], *image), device=work_device, dtype=odt)
            state["weight"] = torch.zeros((1, 1, *image), device=work_device, dtype=odt)
        if state["value"].dtype == odt and int(state["value"].shape[1]) == plan.head_channels[0]:
            plan.sw_run(vol.contiguous(), roi, starts, state["wmap"], state["value"], state["weight"],
                        padding_mode=padding_mode, cval=cval, sw_batch=sw_batch_size, use_graph=cuda_graph)
            return state["value"], state["weight"], state["wmap"]

    def run(batch_starts):
        batch = _extract_starts(vol, batch_starts, roi, padding_mode, cval)
        if batch.device != sw_device:
            batch = batch.to(sw_device, non_blocking=True)
        with torch.no_grad():
            return network(batch)

    def blend(out, batch_starts) -> None:
        if state["value"] is None:
            check_network_output(out, len(batch_starts), None, roi, who)
            cout, odt = int(out.shape[1]), out.dtype
            state["wmap"] = build_sliding_importance_map(roi, mode=mode, device=work_device, dtype=odt)
            state["value"] = torch.zeros((1, cout, *image), device=work_device, dtype=odt)
            state["weight"] = torch.zeros((1, 1, *image), device=work_device, dtype=odt)
        v = state["value"]
        check_network_output(out, len(batch_starts), int(v.shape[1]), roi, who)
        out = out.to(device=work_device, dtype=v.dtype).contiguous()
        _accumulate_batch(out, state["wmap"], v, state["weight"], roi, image, batch_starts)

    rest = starts
    if probe_first and starts:          # the reference probes with the first window alone (window.py:612-639)
        blend(run(starts[:1]), starts[:1])
        rest = starts[1:]
    for b0 in range(0, len(rest), sw_batch_size):
        chunk = rest[b0:b0 + sw_batch_size]
        blend(run(chunk), chunk)
    return state["value"], state["weight"], state["wmap"]


class EagerSlidingWindowEngine:
    """``window.py:530-683`` — ``engine(inputs=[1,C,*spatial], network=fn) -> [1,Cout,*spatial]``.

    Device-resident volumes run as one pass over the eager grid.  HOST volumes (``keep_input_on_cpu`` /
    ``output_device='cpu'``, ``window.py:413-461``) are STREAMED: the z-starts of the grid are processed in groups, the
    input planes of the next group travel host -> pinned staging -> HBM on a side stream while the current group computes,
    finished output planes are normalised and sent back through pinned memory, and the partial planes shared with the
    next group are carried over in place — HBM holds two input slabs and one accumulator slab instead of the volume and
    both accumulators, and the fp sums still associate in grid order (bit-identical to the one-pass result)."""

    def __init__(self, *, roi_size, sw_batch_size: int, overlap, mode: str, padding_mode: str, cval: float,
                 sw_device=None, output_device=None, progress: bool = False, stream_z_starts: int = 0,
                 cuda_graph: bool = True) -> None:
        self.roi_size = tuple(int(v) for v in roi_size)
        self.sw_batch_size = max(1, int(sw_batch_size))
        self.overlap = overlap
        self.mode = _normalize_blending_mode(mode)
        self.padding_mode = padding_mode
        self.cval = float(cval)
        self.sw_device = sw_device
        self.output_device = output_device
        self.progress = bool(progress)
        self.stream_z_starts = int(stream_z_starts)     # z-starts per streamed group (0 = auto: 4)
        self.cuda_graph = bool(cuda_graph)              # native tile loop: replay one captured graph per window batch

    # ---- one pass over `starts` on a device-resident volume, into given (or new) accumulators
    def _run_windows(self, vol, network, starts, image, sw_device, work_device, value=None, weight=None, wmap=None,
                     probe_first=True):
        return run_window_list(vol, network, starts, roi=self.roi_size, image=image, sw_batch_size=self.sw_batch_size,
                               padding_mode=self.padding_mode, cval=self.cval, mode=self.mode, sw_device=sw_device,
                               work_device=work_device, value=value, weight=weight, wmap=wmap, probe_first=probe_first,
                               cuda_graph=self.cuda_graph, who="EagerSlidingWindowEngine")

    def __call__(self, inputs: torch.Tensor, network: Callable[[torch.Tensor], torch.Tensor]) -> torch.Tensor:
        roi = self.roi_size
        nd = len(roi)
        if inputs.dim() < nd + 2:
            raise ValueError("EagerSlidingWindowEngine: inputs must have shape (B, C, *spatial); "
                             f"got shape {tuple(inputs.shape)} for roi_size {roi}.")
        if inputs.shape[0] != 1:
            raise ValueError(f"EagerSlidingWindowEngine currently expects batch size 1; got batch {inputs.shape[0]}.")
        original = tuple(int(v) for v in inputs.shape[-nd:])
        sw_device = torch.device(self.sw_device) if self.sw_device else inputs.device
        output_device = torch.device(self.output_device) if self.output_device else inputs.device
        # the tile loop (crop, blend, normalise) always runs on the GPU
        work_device = sw_device if sw_device.type == "cuda" else (
            inputs.device if inputs.is_cuda else output_device)
        _device_or_raise(work_device)
        grow = [max(0, roi[a] - original[a]) for a in range(nd)]
        if (not inputs.is_cuda) and nd == 3 and not any(grow) and original[0] > roi[0]:
            return self._call_streamed(inputs, network, sw_device, work_device, output_device)
        vol = inputs.to(work_device, non_blocking=True)
        if any(grow):  # constant pad up to the ROI (window.py:583-601): one padded-window gather
            grown = tuple(original[a] + grow[a] for a in range(nd))
            vol = _extract_starts(vol, [(0,) * nd], grown, "constant", self.cval)
        image = tuple(int(v) for v in vol.shape[-nd:])
        starts = _plan(L.GRID_EAGER, image, roi, self.overlap)
        value, weight, _ = self._run_windows(vol, network, starts, image, sw_device, work_device)
        out = normalize_weighted_accumulator(value, weight)
        if any(grow):
            out = out[(slice(None), slice(None)) + tuple(slice(0, original[a]) for a in range(nd))].contiguous()
        return out if out.device == output_device else out.to(output_device)

    # ---- host volume: z-slab streaming with double-buffered H2D on a side stream
    def _call_streamed(self, inputs, network, sw_device, work_device, output_device):
        roi = self.roi_size
        image = tuple(int(v) for v in inputs.shape[-3:])
        cin = int(inputs.shape[1])
        starts = _plan(L.GRID_EAGER, image, roi, self.overlap)
        zs = sorted({s[0] for s in starts})
        g = self.stream_z_starts if self.stream_z_starts > 0 else 4
        groups = [zs[i:i + g] for i in range(0, len(zs), g)]
        depth = max(gz[-1] + roi[0] - gz[0] for gz in groups)
        main = torch.cuda.current_stream(work_device)
        side = torch.cuda.Stream(device=work_device)
        src = inputs if inputs.is_contiguous() else inputs.contiguous()
        pinned_src = src.is_pinned()
        stage = None if pinned_src else [torch.empty((1, cin, depth, image[1], image[2]), dtype=src.dtype).pin_memory()
                                         for _ in range(2)]
        dbuf = [torch.empty((1, cin, depth, image[1], image[2]), device=work_device, dtype=src.dtype) for _ in range(2)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]          # slab i is in HBM
        consumed = [torch.cuda.Event(), torch.cuda.Event()]       # the windows of the group reading dbuf[i] are enqueued+done
        staged = [None, None]                                     # H2D out of stage[i] finished (host may refill it)

        def upload(gi):
            b = gi % 2
            z0, z1 = groups[gi][0], groups[gi][-1] + roi[0]
            if stage is not None:
                if staged[b] is not None:
                    staged[b].synchronize()
                stage[b][:, :, :z1 - z0].copy_(src[:, :, z0:z1])              # pageable -> pinned (host memcpy)
                host = stage[b][:, :, :z1 - z0]
            else:
                host = src[:, :, z0:z1]
            with torch.cuda.stream(side):
                side.wait_event(consumed[b])                                   # do not overwrite a slab still being read
                dbuf[b][:, :, :z1 - z0].copy_(host, non_blocking=True)
                ready[b].record(side)
                if stage is not None:
                    staged[b] = torch.cuda.Event()
                    staged[b].record(side)

        for ev in consumed:
            ev.record(main)
        upload(0)
        out_host = None
        value = weight = wmap = None
        done_ev = None
        out_lo = 0
        for gi, gz in enumerate(groups):
            b = gi % 2
            if gi + 1 < len(groups):
                upload(gi + 1)                 # next slab's H2D overlaps this group's windows
            z0, z1 = gz[0], gz[-1] + roi[0]
            local_image = (z1 - z0, image[1], image[2])
            main.wait_event(ready[b])
            vol = dbuf[b][:, :, :z1 - z0]
            mine = [(s[0] - z0, s[1], s[2]) for s in starts if gz[0] <= s[0] <= gz[-1]]     # grid order preserved
            carry = None
            if value is not None:              # partial planes shared with the previous group go first (same fp order)
                pz0 = groups[gi - 1][0]
                keep = groups[gi - 1][-1] + roi[0] - z0
                if keep > 0:
                    carry = (value[:, :, z0 - pz0:z0 - pz0 + keep], weight[:, :, z0 - pz0:z0 - pz0 + keep], keep)
            if value is not None:              # fresh accumulators for this group's z-extent (+ the carried head)
                nv = torch.zeros((1, value.shape[1], *local_image), device=work_device, dtype=value.dtype)
                nw = torch.zeros((1, 1, *local_image), device=work_device, dtype=value.dtype)
                if carry is not None:
                    nv[:, :, :carry[2]].copy_(carry[0])
                    nw[:, :, :carry[2]].copy_(carry[1])
                value, weight = nv, nw
            value, weight, wmap = self._run_windows(vol, network, mine, local_image, sw_device, work_device, value, weight,
                                                    wmap, probe_first=(gi == 0))
            consumed[b].record(main)
            # planes no later group touches are final: normalise and ship them
            final_hi = image[0] if gi + 1 == len(groups) else groups[gi + 1][0]
            lo, hi = out_lo - z0, final_hi - z0
            if hi > lo:
                part = normalize_weighted_accumulator(value[:, :, lo:hi].contiguous(), weight[:, :, lo:hi].contiguous())
                if output_device.type == "cuda":
                    if out_host is None:
                        out_host = torch.empty((1, value.shape[1], *image), device=output_device, dtype=value.dtype)
                    out_host[:, :, out_lo:final_hi].copy_(part, non_blocking=True)
                else:
                    if out_host is None:
                        out_host = torch.empty((1, value.shape[1], *image), dtype=value.dtype).pin_memory()
                    out_host[:, :, out_lo:final_hi].copy_(part, non_blocking=True)
                    done_ev = torch.cuda.Event()
                    done_ev.record(main)
            out_lo = final_hi
        if done_ev is not None:
            done_ev.synchronize()
        main.wait_stream(side)
        return out_host


def build_sliding_inferer(cfg) -> Optional[EagerSlidingWindowEngine]:
    """``window.py:686-732``."""
    roi = resolve_inferer_roi_size(cfg)
    if roi is None:
        logger.warning("Sliding-window inference disabled: unable to determine ROI size. "
                       "Set inference.window_size or model.output_size in the config.")
        return None
    rt = _resolve_sliding_window_runtime(cfg, roi)
    if resolve_border_mask(cfg, len(roi)):
        logger.warning("inference.sliding_window.border_mask is set but the eager sliding-window engine ignores it; "
                       "use the lazy sliding-window path to apply border masking.")
    return EagerSlidingWindowEngine(roi_size=roi, sw_batch_size=rt["sw_batch_size"], overlap=rt["overlap"],
                                    mode=rt["mode"], padding_mode=rt["padding_mode"], cval=rt["cval"],
                                    sw_device=rt["sw_device"], output_device=rt["output_device"], progress=False)


__all__ = ["EagerSlidingWindowEngine", "apply_border_mask", "build_sliding_accumulator_weight_maps",
           "build_sliding_importance_map", "build_sliding_inferer", "compute_importance_map",
           "compute_scan_interval", "dense_patch_slices", "is_2d_inference_mode",
           "is_distance_transform_blending", "normalize_weighted_accumulator", "resolve_border_mask",
           "resolve_inferer_overlap", "resolve_inferer_roi_size", "resolve_model_output_dtype"]
