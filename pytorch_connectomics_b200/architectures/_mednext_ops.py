"""Functional layer between the MedNeXt modules and the C ABI (``include/pcb200.h``).

Every function here enqueues hand-written sm_100a kernels on ``torch.cuda.current_stream()``; torch
is used for allocation and for autograd bookkeeping only.  Activations are channels-last bf16
``[N, D, H, W, C]``.  ``torch.autograd.Function`` wrappers store, per block, the block input and the
depthwise-conv output (+ its GroupNorm statistics); the expanded ``[V, r*C]`` tensor is never stored —
backward recomputes it on the tensor cores (the reference's ``outside_block`` checkpointing recomputes
whole blocks; this is the same trade at a finer grain).
"""

from __future__ import annotations

import ctypes
import weakref
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .. import _lib as L

_BF16 = torch.bfloat16

# ----------------------------------------------------------------------------- weight repacking cache
_CACHE: Dict[Tuple[int, str], tuple] = {}


def _pack_dw(w):      # [C,1,k,k,k] -> [k^3, C] f32 (tap-major)
    c = w.shape[0]
    return w.reshape(c, -1).t().contiguous().float()


def _pack_pw(w):      # Conv3d 1x1 [O,I,1,1,1] -> [O,I] bf16
    return w.reshape(w.shape[0], w.shape[1]).contiguous().to(_BF16)


def _pack_pw_t(w):    # ConvTranspose3d 1x1 [I,O,1,1,1] -> [O,I] bf16
    return w.reshape(w.shape[0], w.shape[1]).t().contiguous().to(_BF16)


def _pack_f32(w):
    return w.reshape(-1).contiguous().float()


def _pack_head(w):    # ConvTranspose3d 1x1 [C,ncls,1,1,1] -> [C,ncls] f32
    return w.reshape(w.shape[0], w.shape[1]).contiguous().float()


def _pack_head_conv(w):  # Conv3d 1x1 [ncls,C,1,1,1] -> [C,ncls] f32
    return w.reshape(w.shape[0], w.shape[1]).t().contiguous().float()


_PACKERS = {"dw": _pack_dw, "pw": _pack_pw, "pw_t": _pack_pw_t, "f32": _pack_f32, "head": _pack_head,
            "head_conv": _pack_head_conv}


def packed(p: torch.Tensor, kind: str) -> torch.Tensor:
    """Kernel-layout copy of a parameter, refreshed when the parameter changes (optimizer step,
    load_state_dict, .to(device))."""
    key = (id(p), kind)
    ent = _CACHE.get(key)
    ver = (p._version, L.PARAM_EPOCH[0])
    if ent is not None and ent[0] == ver and ent[1]() is p and ent[2] == p.data_ptr():
        return ent[3]
    with torch.no_grad():
        t = _PACKERS[kind](p.detach())
    if len(_CACHE) > 8192:
        for k in [k for k, e in _CACHE.items() if e[1]() is None]:
            del _CACHE[k]
    _CACHE[key] = (ver, weakref.ref(p), p.data_ptr(), t)
    return t


# ----------------------------------------------------------------------------- helpers
def as_channels_last(x: torch.Tensor) -> torch.Tensor:
    """Accept [N,D,H,W,C] bf16 (internal) or an NCDHW tensor/view; return contiguous [N,D,H,W,C] bf16."""
    if getattr(x, "_pcb_cl", False):
        return x
    if x.dim() != 5:
        raise ValueError(f"expected a 5-D tensor, got shape {tuple(x.shape)}")
    L.require_device(x, "channels-last conversion")
    y = x.permute(0, 2, 3, 4, 1)
    if y.dtype != _BF16:
        y = y.to(_BF16)
    y = y.contiguous()
    y._pcb_cl = True
    return y


def as_channels_last_2d(x: torch.Tensor) -> torch.Tensor:
    """2-D networks run as depth-1 volumes: accept the internal [N,1,H,W,C] tensor or an NCHW tensor (lifted to N,C,1,H,W)."""
    if getattr(x, "_pcb_cl", False):
        return x
    if x.dim() != 4:
        raise ValueError(f"expected a 4-D (N, C, H, W) tensor, got shape {tuple(x.shape)}")
    return as_channels_last(x.unsqueeze(2))


def _mark(t: torch.Tensor) -> torch.Tensor:
    t._pcb_cl = True
    return t


def _out_size(mode: int, k: int, size: Sequence[int]) -> Tuple[List[int], List[int]]:
    """(dw-conv output size, block output size) for an input spatial size."""
    p = k // 2
    if mode == L.DW_SAME:
        return list(size), list(size)
    if mode == L.DW_DOWN:
        o = [(s + 2 * p - k) // 2 + 1 for s in size]
        return o, o
    y = [(s - 1) * 2 - 2 * p + k for s in size]      # ConvTranspose, 2s-1 for k=3,5,7 with pad k//2
    return y, [v + 1 for v in y]


# ----------------------------------------------------------------------------- raw forward launches
def stem_forward(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    x = x.contiguous()
    n, cin = int(x.shape[0]), int(x.shape[1])
    c = int(w.shape[0])
    out = torch.empty((n, *x.shape[2:], c), device=x.device, dtype=_BF16)
    nvox = int(x.shape[2] * x.shape[3] * x.shape[4])
    L.check(L.lib().pcb_stem_fwd(L.ptr(x), L.dtype_code(x.dtype), L.ptr(packed(w, "f32")), L.ptr(packed(b, "f32")),
                                 L.ptr(out), ctypes.c_int64(n), ctypes.c_int64(cin), ctypes.c_int64(c),
                                 ctypes.c_int64(nvox), L.stream_ptr(x.device)), "pcb_stem_fwd")
    return _mark(out)


_IDENT: Dict[tuple, tuple] = {}


def identity_groupnorm(n: int, c: int, vy: int, device) -> tuple:
    """(stats [n,2,c] f64, ones [c] f32, zeros [c] f32) that make the kernels' GroupNorm-apply the identity
    (sum 0, sum of squares V*(1-eps) -> mean 0, rstd 1): used when the block's norm is the channels-first
    LayerNorm, which runs as its own kernel in front of the fused MLP."""
    key = (str(device), n, c, vy)
    ent = _IDENT.get(key)
    if ent is None:
        stats = torch.zeros((n, 2, c), device=device, dtype=torch.float64)
        stats[:, 1] = float(vy) * (1.0 - 1e-5)
        ent = (stats, torch.ones(c, device=device, dtype=torch.float32), torch.zeros(c, device=device, dtype=torch.float32))
        if len(_IDENT) > 256:
            _IDENT.clear()
        _IDENT[key] = ent
    return ent


def layernorm_forward(y: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """upstream blocks.py::LayerNorm(channels_first) on a channels-last bf16 tensor (``pcb_layernorm_fwd``)."""
    c = int(y.shape[-1])
    rows = y.numel() // c
    out = torch.empty_like(y)
    st = L.lib().pcb_layernorm_fwd(L.ptr(y), L.ptr(packed(weight, "f32")), L.ptr(packed(bias, "f32")), L.ptr(out),
                                   ctypes.c_int64(c), ctypes.c_int64(rows), L.stream_ptr(y.device))
    if st == -3:
        raise NotImplementedError(L.lib().pcb_last_error().decode())
    L.check(st, "pcb_layernorm_fwd")
    return out


def block_forward(x: torch.Tensor, skip: Optional[torch.Tensor], params: List[torch.Tensor], mode: int, k: int,
                  do_res: bool, has_rc: bool, norm: str = "group"):
    """Returns (out, y, stats).  x: [N,D,H,W,C] bf16."""
    w1, b1, gamma, beta, w2, b2, w3, b3 = params[:8]
    n, size, c = int(x.shape[0]), [int(s) for s in x.shape[1:4]], int(x.shape[4])
    h, co = int(w2.shape[0]), int(w3.shape[0])
    ysize, osize = _out_size(mode, k, size)
    dev = x.device
    st = L.stream_ptr(dev)
    lib = L.lib()
    y = torch.empty((n, *ysize, c), device=dev, dtype=_BF16)
    stats = torch.zeros((n, 2, c), device=dev, dtype=torch.float64)
    vy = ysize[0] * ysize[1] * ysize[2]
    with L.prof(f"dwconv_fwd:m{mode}C{c}V{vy}"):
        L.check(lib.pcb_dwconv_fwd(L.ptr(x), L.ptr(packed(w1, "dw")), L.ptr(packed(b1, "f32")), L.ptr(y), L.ptr(stats),
                                   ctypes.c_int64(n), L.i64x(size), ctypes.c_int64(c), ctypes.c_int(k),
                                   ctypes.c_int(mode), st), "pcb_dwconv_fwd")
    y_in, gam_t, bet_t = y, packed(gamma, "f32"), packed(beta, "f32")
    if norm == "layer":
        with L.prof(f"layernorm_fwd:C{c}V{vy}"):
            y_in = layernorm_forward(y, gamma, beta)
        stats_in, gam_t, bet_t = identity_groupnorm(n, c, vy, dev)
    else:
        stats_in = stats
    out = torch.empty((n, *osize, co), device=dev, dtype=_BF16)
    res = None
    if mode == L.DW_SAME and do_res:
        res = x
    elif mode == L.DW_UP and skip is not None:
        res = skip
        if tuple(skip.shape) != tuple(out.shape):
            raise ValueError(f"skip shape {tuple(skip.shape)} != up-block output shape {tuple(out.shape)}")
    wr = br = None
    cr = 0
    if has_rc:
        wr = packed(params[8], "pw_t" if mode == L.DW_UP else "pw")
        br = packed(params[9], "f32")
        cr = c
    vo = osize[0] * osize[1] * osize[2]
    ws = int(lib.pcb_mlp_fwd_deep_workspace(ctypes.c_int64(n), L.i64x(osize), ctypes.c_int64(c), ctypes.c_int64(h),
                                            ctypes.c_int64(co), ctypes.c_int64(cr)))
    if ws > 0:      # deep levels: two column-split GEMM launches through an HBM/L2-resident expanded activation
        hact = torch.empty((n, *osize, h), device=dev, dtype=_BF16)
        with L.prof(f"mlp_fwd_deep:m{mode}C{c}H{h}Co{co}V{vo}"):
            L.check(lib.pcb_mlp_fwd_deep(L.ptr(y_in), L.ptr(stats_in), L.ptr(gam_t), L.ptr(bet_t),
                                         L.ptr(packed(w2, "pw")), L.ptr(packed(b2, "f32")), L.ptr(packed(w3, "pw")),
                                         L.ptr(packed(b3, "f32")), L.ptr(res), L.ptr(x if has_rc else None), L.ptr(wr),
                                         L.ptr(br), L.ptr(out), L.ptr(hact), ctypes.c_int64(n), L.i64x(osize), L.i64x(size),
                                         ctypes.c_int64(c), ctypes.c_int64(h), ctypes.c_int64(co), ctypes.c_int64(cr),
                                         ctypes.c_int(mode), st), "pcb_mlp_fwd_deep")
        return _mark(out), y, stats
    with L.prof(f"mlp_fwd:m{mode}C{c}H{h}Co{co}V{vo}"):
        L.check(lib.pcb_mlp_fwd(L.ptr(y_in), L.ptr(stats_in), L.ptr(gam_t), L.ptr(bet_t),
                                L.ptr(packed(w2, "pw")), L.ptr(packed(b2, "f32")), L.ptr(packed(w3, "pw")),
                                L.ptr(packed(b3, "f32")), L.ptr(res), L.ptr(x if has_rc else None), L.ptr(wr), L.ptr(br),
                                L.ptr(out), ctypes.c_int64(n), L.i64x(osize), L.i64x(size), ctypes.c_int64(c),
                                ctypes.c_int64(h), ctypes.c_int64(co), ctypes.c_int64(cr), ctypes.c_int(mode), st),
                "pcb_mlp_fwd")
    return _mark(out), y, stats


def dwconv_forward(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, mode: int, k: int) -> torch.Tensor:
    """The depthwise stencil ALONE (``pcb_dwconv_fwd``: same / stride-2 / transposed stride-2), for blocks that are composed
    kernel by kernel (GRN) instead of running the fused block."""
    n, size, c = int(x.shape[0]), [int(s) for s in x.shape[1:4]], int(x.shape[4])
    ysize, _ = _out_size(mode, k, size)
    y = torch.empty((n, *ysize, c), device=x.device, dtype=_BF16)
    stats = torch.zeros((n, 2, c), device=x.device, dtype=torch.float64)
    L.check(L.lib().pcb_dwconv_fwd(L.ptr(x), L.ptr(packed(w1, "dw")), L.ptr(packed(b1, "f32")), L.ptr(y), L.ptr(stats),
                                   ctypes.c_int64(n), L.i64x(size), ctypes.c_int64(c), ctypes.c_int(k), ctypes.c_int(mode),
                                   L.stream_ptr(x.device)), "pcb_dwconv_fwd")
    return _mark(y)


def head_forward(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, out_dtype: torch.dtype,
                 conv_layout: bool = False) -> torch.Tensor:
    n, c = int(x.shape[0]), int(x.shape[4])
    wk = packed(w, "head_conv" if conv_layout else "head")
    ncls = int(wk.shape[1])
    out = torch.empty((n, ncls, *x.shape[1:4]), device=x.device, dtype=out_dtype)
    nvox = int(x.shape[1] * x.shape[2] * x.shape[3])
    L.check(L.lib().pcb_head_fwd(L.ptr(x), L.ptr(wk), L.ptr(packed(b, "f32")), L.ptr(out),
                                 ctypes.c_int(L.dtype_code(out_dtype)), ctypes.c_int64(n), ctypes.c_int64(c),
                                 ctypes.c_int64(ncls), ctypes.c_int64(nvox), L.stream_ptr(x.device)), "pcb_head_fwd")
    return out


# ----------------------------------------------------------------------------- autograd wrappers
def _needs_grad(*ts) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts)


def stem_apply(x, w, b):
    L.require_device(x, "MedNeXt stem")
    if _needs_grad(x, w, b):
        from . import _mednext_bwd as B
        return _mark(B.StemFn.apply(x, w, b))
    return stem_forward(x, w, b)


def block_apply(x, skip, params, mode, k, do_res, has_rc, norm="group"):
    x = as_channels_last(x)
    if skip is not None:
        skip = as_channels_last(skip)
    if _needs_grad(x, skip, *params):
        from . import _mednext_bwd as B
        return _mark(B.BlockFn.apply(x, skip, mode, k, do_res, has_rc, norm, *params))
    return block_forward(x, skip, params, mode, k, do_res, has_rc, norm)[0]


def dwconv_apply(x, w1, b1, mode, k):
    x = as_channels_last(x)
    if _needs_grad(x, w1, b1):
        from . import _mednext_bwd as B
        return _mark(B.DwConvFn.apply(x, w1, b1, mode, k))
    return dwconv_forward(x, w1, b1, mode, k)


_ONE: Dict[str, torch.Tensor] = {}


def norm_apply(y, gamma, beta, norm):
    """The block's normalisation ALONE on channels-last bf16: ``GroupNorm(C groups)`` through the statistics / per-channel
    affine kernels (``GroupNormActFn`` of the dense-conv family with a PReLU slope of 1, i.e. no activation) or the
    channels-first LayerNorm kernel — with their backward kernels under autograd."""
    y = as_channels_last(y)
    if norm == "layer":
        if _needs_grad(y, gamma, beta):
            from . import _mednext_bwd as B
            return _mark(B.LayerNormFn.apply(y, gamma, beta))
        return _mark(layernorm_forward(y, gamma, beta))
    from .monai_unet import GroupNormActFn
    one = _ONE.get(str(y.device))
    if one is None:
        one = _ONE[str(y.device)] = torch.ones(1, device=y.device, dtype=torch.float32)
    c = int(y.shape[4])
    return _mark(GroupNormActFn.apply(y, gamma, beta, one, c, 1e-5, c))


def head_apply(x, w, b, out_dtype, conv_layout=False):
    x = as_channels_last(x)
    if _needs_grad(x, w, b):
        from . import _mednext_bwd as B
        return B.HeadFn.apply(x, w, b, out_dtype, conv_layout)
    return head_forward(x, w, b, out_dtype, conv_layout)


def pointwise_forward(x: torch.Tensor, w_packed: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    """1x1 conv on channels-last bf16: out[..., :] = x[..., :] @ w_packed[N,K]^T + b   (tcgen05 GEMM)."""
    n, size, k = int(x.shape[0]), [int(s) for s in x.shape[1:4]], int(x.shape[4])
    nw = int(w_packed.shape[0])
    if k % 16 or nw % 16:
        raise ValueError(f"pcb200: 1x1 projection needs channel counts that are multiples of 16 (got {k}->{nw})")
    out = torch.empty((n, *size, nw), device=x.device, dtype=_BF16)
    L.check(L.lib().pcb_pw_fwd(L.ptr(x), L.ptr(w_packed), L.ptr(b), L.ptr(out), ctypes.c_int64(n), L.i64x(size), 0,
                               L.i64x(size), ctypes.c_int64(k), ctypes.c_int64(nw), L.stream_ptr(x.device)), "pcb_pw_fwd")
    return _mark(out)


def pointwise_apply(x, w, b):
    """MedNeXtTaskHead.input_projection (``mednext_models.py:169-173``): Conv3d(C, hidden, 1)."""
    x = as_channels_last(x)
    if _needs_grad(x, w, b):
        from . import _mednext_bwd as B
        return _mark(B.PointwiseFn.apply(x, w, b))
    return pointwise_forward(x, packed(w, "pw"), packed(b, "f32"))
