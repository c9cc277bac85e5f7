"""TEST INFRASTRUCTURE — CPU stand-ins for the kernel-calling entry points of ``architectures/_mednext_ops.py`` (stem, block,
head, pointwise projection, and the single-kernel pieces the GRN composition uses), written in plain differentiable torch from
the contract each kernel states: channels-last ``[N, D, H, W, C]`` activations in and out, parameters in the UPSTREAM 3-D
shapes.  They let the HOST logic above the kernels (the 2-D lift of ``dim="2d"``, the GRN composition, wrappers and heads) run
in the CPU suite in fp32 and be compared with the oracle network, values and gradients.  Never the thing under test, never
reachable from the product package: the GPU tests run the same modules against the real kernels."""

from __future__ import annotations

import torch
import torch.nn.functional as F

from pytorch_connectomics_b200 import _lib as L
from pytorch_connectomics_b200.architectures import _mednext_ops as ops


def _cf(x):      # channels-last -> channels-first
    return x.permute(0, 4, 1, 2, 3)


def _cl(x):      # channels-first -> marked channels-last
    return ops._mark(x.permute(0, 2, 3, 4, 1).contiguous())


def _flat(w):    # any 1x1 weight [A, B, 1, ...] -> [A, B]
    return w.reshape(w.shape[0], w.shape[1])


def _norm(y, gamma, beta, norm):
    if norm == "group":
        return F.group_norm(y, int(y.shape[1]), gamma, beta, 1e-5)
    u = y.mean(1, keepdim=True)
    s = (y - u).pow(2).mean(1, keepdim=True)
    return gamma.view(1, -1, 1, 1, 1) * ((y - u) / torch.sqrt(s + 1e-5)) + beta.view(1, -1, 1, 1, 1)


def dwconv_apply(x, w1, b1, mode, k):
    xc, c, p = _cf(x), int(x.shape[4]), k // 2
    if mode == L.DW_SAME:
        y = F.conv3d(xc, w1, b1, padding=p, groups=c)
    elif mode == L.DW_DOWN:
        y = F.conv3d(xc, w1, b1, stride=2, padding=p, groups=c)
    else:
        y = F.conv_transpose3d(xc, w1, b1, stride=2, padding=p, groups=c)
    return _cl(y)


def norm_apply(y, gamma, beta, norm):
    return _cl(_norm(_cf(y), gamma, beta, norm))


def pointwise_apply(x, w, b):
    y = torch.einsum("ndhwc,oc->ndhwo", x, _flat(w))
    return ops._mark(y + b if b is not None else y)


def block_apply(x, skip, params, mode, k, do_res, has_rc, norm="group"):
    w1, b1, gamma, beta, w2, b2, w3, b3 = params[:8]
    xc = _cf(x)
    y = _cf(dwconv_apply(x, w1, b1, mode, k))
    a = _norm(y, gamma, beta, norm)
    h = F.gelu(torch.einsum("nchwd,oc->nohwd", a, _flat(w2)) + b2.view(1, -1, 1, 1, 1))
    o = torch.einsum("nchwd,oc->nohwd", h, _flat(w3)) + b3.view(1, -1, 1, 1, 1)
    if mode == L.DW_SAME and do_res:
        o = o + xc
    if has_rc:
        wr, br = params[8], params[9]
        if mode == L.DW_DOWN:
            o = o + F.conv3d(xc, _flat(wr)[:, :, None, None, None], br, stride=2)
        else:
            o = o + F.conv_transpose3d(xc, _flat(wr)[:, :, None, None, None], br, stride=2)
    if mode == L.DW_UP:
        o = F.pad(o, (1, 0, 1, 0, 1, 0))
        if skip is not None:
            o = o + _cf(skip)
    return _cl(o)


def stem_apply(x, w, b):
    return _cl(F.conv3d(x, _flat(w)[:, :, None, None, None], b))


def head_apply(x, w, b, out_dtype, conv_layout=False):
    wk = _flat(w) if conv_layout else _flat(w).t()          # -> [ncls, C]
    return (torch.einsum("ndhwc,oc->nodhw", x, wk) + b.view(1, -1, 1, 1, 1)).to(out_dtype)


def install(monkeypatch) -> None:
    """route the MedNeXt entry points to the stand-ins and keep activations fp32 for the duration of one test"""
    monkeypatch.setattr(L, "require_device", lambda t, what: None)
    monkeypatch.setattr(ops, "_BF16", torch.float32)
    for name in ("block_apply", "stem_apply", "head_apply", "pointwise_apply", "dwconv_apply", "norm_apply"):
        monkeypatch.setattr(ops, name, globals()[name], raising=False)
