"""MONAI ``UNet`` on the B200 engine — drop-in for what ``build_monai_unet``
(``connectomics/models/architectures/monai_models.py:197-250``) gets from ``monai.networks.nets.UNet``
(BASELINE config 1, ``tutorials/minimal.yaml``).

Module / child names follow MONAI (``Convolution`` = Sequential{conv, adn{N,D,A}}, ``ResidualUnit`` =
{conv{unit0..}, residual}, ``SkipConnection.submodule``), so ``state_dict`` keys are the MONAI ones
(``model.0.conv.unit0.conv.weight`` …) and reference checkpoints load unchanged.  The torch children are
parameter/buffer containers; the arithmetic runs in ``csrc/dense_conv.cu`` (implicit-GEMM conv on tcgen05,
BatchNorm+PReLU, their backward) and ``csrc/mednext_bwd.cu`` (split-K weight-gradient GEMM) on channels-last
bf16 activations whose channel counts are zero-padded to multiples of 16.  Not yet on hand-written kernels
(plain tensor plumbing for now): the residual ``+`` and the skip ``cat``.
Supported: 3-D, ``norm="batch"`` / ``"group"`` / ``"instance"``, ``dropout`` (identity at inference; in training the mask is
applied behind the fused norm+PReLU kernel, which equals MONAI's norm -> dropout -> PReLU for the same mask),
``upsample_mode="deconv"`` / ``"nontrainable"`` — anything else raises.
"""

from __future__ import annotations

import ctypes
from typing import List, Sequence

import torch
import torch.nn as nn

from .. import _lib as L
from . import _mednext_ops as ops
from .base import ConnectomicsModel
from .registry import register_architecture

_BF16 = torch.bfloat16


def _pad16(c: int) -> int:
    return (int(c) + 15) // 16 * 16


def _pad_mat(w: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    """[..., r, c] -> zero-padded [..., rows, cols] bf16 contiguous."""
    out = torch.zeros(w.shape[:-2] + (rows, cols), device=w.device, dtype=_BF16)
    out[..., : w.shape[-2], : w.shape[-1]] = w.to(_BF16)
    return out.contiguous()


def _pack(kind: str):
    def conv_fwd(w):       # Conv3d [Co,Ci,k,k,k] -> [k^3][Co][Ci]
        co, ci = w.shape[:2]
        return _pad_mat(w.permute(2, 3, 4, 0, 1).reshape(-1, co, ci), _pad16(co), _pad16(ci))

    def conv_dgrad_s1(w):  # flipped taps, [k^3][Ci][Co]
        co, ci = w.shape[:2]
        return _pad_mat(w.flip(2, 3, 4).permute(2, 3, 4, 1, 0).reshape(-1, ci, co), _pad16(ci), _pad16(co))

    def conv_dgrad_s2(w):  # unflipped, [k^3][Ci][Co]  (used by the transposed gather)
        co, ci = w.shape[:2]
        return _pad_mat(w.permute(2, 3, 4, 1, 0).reshape(-1, ci, co), _pad16(ci), _pad16(co))

    def convt_fwd(w):      # ConvTranspose3d [Ci,Co,k,k,k] -> [k^3][Co][Ci]
        ci, co = w.shape[:2]
        return _pad_mat(w.permute(2, 3, 4, 1, 0).reshape(-1, co, ci), _pad16(co), _pad16(ci))

    def convt_dgrad(w):    # [k^3][Ci][Co]
        ci, co = w.shape[:2]
        return _pad_mat(w.permute(2, 3, 4, 0, 1).reshape(-1, ci, co), _pad16(ci), _pad16(co))

    def vec16(v):
        out = torch.zeros(_pad16(v.numel()), device=v.device, dtype=torch.float32)
        out[: v.numel()] = v.reshape(-1).float()
        return out

    def ones16(v):
        out = torch.ones(_pad16(v.numel()), device=v.device, dtype=torch.float32)
        out[: v.numel()] = v.reshape(-1).float()
        return out

    return {"conv_fwd": conv_fwd, "conv_dgrad_s1": conv_dgrad_s1, "conv_dgrad_s2": conv_dgrad_s2, "convt_fwd": convt_fwd,
            "convt_dgrad": convt_dgrad, "vec16": vec16, "ones16": ones16}[kind]


for _k in ("conv_fwd", "conv_dgrad_s1", "conv_dgrad_s2", "convt_fwd", "convt_dgrad", "vec16", "ones16"):
    ops._PACKERS.setdefault(_k, _pack(_k))


def _conv_launch(x, w_packed, bias, out_size, ci_p, co_p, k, stride, pad, transposed):
    n = int(x.shape[0])
    in_size = [int(s) for s in x.shape[1:4]]
    out = torch.empty((n, *out_size, co_p), device=x.device, dtype=_BF16)
    L.check(L.lib().pcb_conv_fwd(L.ptr(x), L.ptr(w_packed), L.ptr(bias), L.ptr(out), ctypes.c_int64(n), L.i64x(in_size),
                                 L.i64x(out_size), ctypes.c_int64(ci_p), ctypes.c_int64(co_p), k, stride, pad,
                                 1 if transposed else 0, L.stream_ptr(x.device)), "pcb_conv_fwd")
    return out


class ConvFn(torch.autograd.Function):
    """Conv3d / ConvTranspose3d (+bias) on padded channels-last bf16 through the implicit-GEMM kernel."""

    @staticmethod
    def forward(ctx, x, weight, bias, k, stride, pad, transposed):
        in_size = [int(s) for s in x.shape[1:4]]
        if transposed:
            ci, co = int(weight.shape[0]), int(weight.shape[1])
            out_size = [(s - 1) * stride - 2 * pad + k + (stride - 1) for s in in_size]   # output_padding = stride-1
            wp = ops.packed(weight, "convt_fwd")
        else:
            co, ci = int(weight.shape[0]), int(weight.shape[1])
            out_size = [(s + 2 * pad - k) // stride + 1 for s in in_size]
            wp = ops.packed(weight, "conv_fwd")
        if int(x.shape[4]) != _pad16(ci):
            raise ValueError(f"conv expects {_pad16(ci)} (padded) input channels, got {int(x.shape[4])}")
        bp = ops.packed(bias, "vec16") if bias is not None else None
        out = _conv_launch(x, wp, bp, out_size, _pad16(ci), _pad16(co), k, stride, pad, transposed)
        ctx.save_for_backward(x, weight)
        ctx.cfg = (k, stride, pad, transposed, bias is not None, ci, co, in_size, out_size)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        k, stride, pad, transposed, has_bias, ci, co, in_size, out_size = ctx.cfg
        dy = dy.contiguous()
        n, cip, cop = int(x.shape[0]), _pad16(ci), _pad16(co)
        lib, st, dev = L.lib(), L.stream_ptr(x.device), x.device
        dx = None
        if ctx.needs_input_grad[0]:
            if transposed:      # gradient of a transposed conv is the strided conv
                dx = _conv_launch(dy, ops.packed(weight, "convt_dgrad"), None, in_size, cop, cip, k, stride, pad, False)
            elif stride == 1:   # same-size conv with flipped taps
                dx = _conv_launch(dy, ops.packed(weight, "conv_dgrad_s1"), None, in_size, cop, cip, k, 1, pad, False)
            else:               # gradient of a strided conv is the transposed gather
                dx = _conv_launch(dy, ops.packed(weight, "conv_dgrad_s2"), None, in_size, cop, cip, k, stride, pad, True)
        # weight gradient: one split-K GEMM per tap, D[co, ci] = sum_o dy[o,co] * x[src(o,tap), ci]
        k3 = k * k * k
        if transposed:
            dwp = torch.zeros((cip, cop, k3), device=dev, dtype=torch.float32)
            ldm, ldn = k3, cop * k3
        else:
            dwp = torch.zeros((cop, cip, k3), device=dev, dtype=torch.float32)
            ldm, ldn = cip * k3, k3
        dbp = torch.zeros((cop,), device=dev, dtype=torch.float32)
        nfl = int(lib.pcb_tn_workspace_floats(ctypes.c_int64(cop), ctypes.c_int64(cip), 1, ctypes.c_int64(n), L.i64x(out_size)))
        ws = torch.empty(nfl, device=dev, dtype=torch.float32)
        for t in range(k3):
            tap = (ctypes.c_int * 3)(t // (k * k), (t // k) % k, t % k)
            dst = ctypes.c_void_p(dwp.data_ptr() + 4 * t)
            L.check(lib.pcb_conv_wgrad_tap(L.ptr(dy), L.ptr(x), L.ptr(ws), dst, ctypes.c_int64(ldm), ctypes.c_int64(ldn),
                                           L.ptr(dbp) if t == 0 else None, ctypes.c_int64(n), L.i64x(out_size), L.i64x(in_size),
                                           ctypes.c_int64(cop), ctypes.c_int64(cip), tap, stride, pad, 1 if transposed else 0,
                                           st), "pcb_conv_wgrad_tap")
        if transposed:
            dw = dwp[:ci, :co].reshape(ci, co, k, k, k)
        else:
            dw = dwp[:co, :ci].reshape(co, ci, k, k, k)
        return dx, dw.contiguous(), (dbp[:co].clone() if has_bias else None), None, None, None, None


class BnActFn(torch.autograd.Function):
    """ADN "NDA" with BatchNorm + PReLU (dropout 0): y = PReLU(BN(x)); batch statistics in training mode."""

    @staticmethod
    def forward(ctx, x, gamma, beta, slope, running_mean, running_var, training, momentum, eps, c_real):
        n, cp = int(x.shape[0]), int(x.shape[4])
        v = int(x.shape[1] * x.shape[2] * x.shape[3])
        rows = n * v
        lib, st = L.lib(), L.stream_ptr(x.device)
        g16, b16 = ops.packed(gamma, "ones16"), ops.packed(beta, "vec16")
        if training:
            stats = torch.zeros((2, cp), device=x.device, dtype=torch.float64)
            L.check(lib.pcb_channel_stats(L.ptr(x), L.ptr(stats), ctypes.c_int64(cp), ctypes.c_int64(rows), st), "pcb_channel_stats")
            mean64 = stats[0] / rows
            var64 = (stats[1] / rows - mean64 * mean64).clamp_min(0.0)
            with torch.no_grad():   # running statistics as nn.BatchNorm3d: momentum update with the unbiased variance
                running_mean.mul_(1 - momentum).add_(mean64[:c_real].float(), alpha=momentum)
                running_var.mul_(1 - momentum).add_((var64[:c_real] * (rows / max(rows - 1, 1))).float(), alpha=momentum)
        else:
            stats = None
            mean64 = torch.zeros(cp, device=x.device, dtype=torch.float64)
            var64 = torch.ones(cp, device=x.device, dtype=torch.float64)
            mean64[:c_real] = running_mean.double()
            var64[:c_real] = running_var.double()
        rstd = (1.0 / torch.sqrt(var64 + eps)).float()
        mean = mean64.float()
        scale = (g16 * rstd).contiguous()
        shift = (b16 - mean * scale).contiguous()
        slope_f = slope.detach().reshape(-1)[:1].float().contiguous()
        out = torch.empty_like(x)
        L.check(lib.pcb_bn_act_fwd(L.ptr(x), L.ptr(scale), L.ptr(shift), L.ptr(slope_f), L.ptr(out), ctypes.c_int64(cp),
                                   ctypes.c_int64(rows), st), "pcb_bn_act_fwd")
        ctx.save_for_backward(x, scale, shift, mean, rstd, slope_f, g16, stats if stats is not None else mean64)
        ctx.cfg = (training, c_real, n, v, cp)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, scale, shift, mean, rstd, slope_f, g16, stats = ctx.saved_tensors
        training, c_real, n, v, cp = ctx.cfg
        dy = dy.contiguous()
        lib, st = L.lib(), L.stream_ptr(x.device)
        rows = n * v
        red = torch.zeros(2 * cp + 1, device=x.device, dtype=torch.float64)
        dz = torch.empty_like(x)
        L.check(lib.pcb_bn_act_bwd(L.ptr(dy), L.ptr(x), L.ptr(scale), L.ptr(shift), L.ptr(mean), L.ptr(rstd), L.ptr(slope_f),
                                   L.ptr(dz), L.ptr(red), ctypes.c_int64(cp), ctypes.c_int64(rows), st), "pcb_bn_act_bwd")
        if training:
            dx = torch.empty_like(x)
            stats_rep = stats.reshape(1, 2, cp).expand(n, 2, cp).contiguous()
            gst_rep = red[: 2 * cp].reshape(1, 2, cp).expand(n, 2, cp).contiguous()
            dsum = torch.zeros(cp, device=x.device, dtype=torch.float64)
            L.check(lib.pcb_bn_bwd(L.ptr(dz), L.ptr(x), L.ptr(stats_rep), L.ptr(gst_rep), L.ptr(g16), L.ptr(dx), L.ptr(dsum),
                                   ctypes.c_int64(n), ctypes.c_int64(cp), ctypes.c_int64(v), st), "pcb_bn_bwd")
        else:               # eval mode: BatchNorm is a fixed affine
            dx = (dz.float() * scale).to(_BF16)
        dgamma = red[cp: cp + c_real].float()
        dbeta = red[:c_real].float()
        dslope = red[2 * cp].float().reshape(1)
        return dx, dgamma, dbeta, dslope, None, None, None, None, None, None


# ----------------------------------------------------------------------------- GroupNorm + PReLU (norm=("group", {...}))
# thin wrappers of the four kernels the function below composes (the CPU suite swaps them for torch stand-ins to check the
# composition's arithmetic; the kernels themselves are the ones BatchNorm runs through)
def _k_channel_stats(x, stats):
    cp = int(x.shape[-1])
    L.check(L.lib().pcb_channel_stats(L.ptr(x), L.ptr(stats), ctypes.c_int64(cp), ctypes.c_int64(x.numel() // cp),
                                      L.stream_ptr(x.device)), "pcb_channel_stats")


def _k_bn_act_fwd(x, scale, shift, slope, out):
    cp = int(x.shape[-1])
    L.check(L.lib().pcb_bn_act_fwd(L.ptr(x), L.ptr(scale), L.ptr(shift), L.ptr(slope), L.ptr(out), ctypes.c_int64(cp),
                                   ctypes.c_int64(x.numel() // cp), L.stream_ptr(x.device)), "pcb_bn_act_fwd")


def _k_bn_act_bwd(dy, x, scale, shift, mean, rstd, slope, dz, red):
    cp = int(x.shape[-1])
    L.check(L.lib().pcb_bn_act_bwd(L.ptr(dy), L.ptr(x), L.ptr(scale), L.ptr(shift), L.ptr(mean), L.ptr(rstd), L.ptr(slope),
                                   L.ptr(dz), L.ptr(red), ctypes.c_int64(cp), ctypes.c_int64(x.numel() // cp),
                                   L.stream_ptr(x.device)), "pcb_bn_act_bwd")


def _k_gn_bwd(g, x, stats, gstats, gamma, dx, dsum):
    n, cp = int(x.shape[0]), int(x.shape[-1])
    L.check(L.lib().pcb_gn_bwd(L.ptr(g), L.ptr(x), L.ptr(stats), L.ptr(gstats), L.ptr(gamma), L.ptr(dx), L.ptr(dsum),
                               ctypes.c_int64(n), ctypes.c_int64(cp), ctypes.c_int64(x.numel() // (n * cp)),
                               L.stream_ptr(x.device)), "pcb_gn_bwd")


class GroupNormActFn(torch.autograd.Function):
    """ADN "NDA" with ``GroupNorm(num_groups)`` + PReLU on padded channels-last bf16.  Statistics are per (sample, group):
    the per-channel sums of ``pcb_channel_stats`` are pooled over each group's channels on the host side (tiny f64 tensors),
    which turns the normalisation into a per-(sample, channel) affine for the BatchNorm apply kernel.  Backward: with
    ``dz = dy * PReLU'(z)`` and the per-channel sums ``S1 = sum dz``, ``S2 = sum dz * xhat`` (``pcb_bn_act_bwd``),
    ``dgamma = S2``, ``dbeta = S1`` and ``dx = rstd_g * (gamma_c * dz - M1_g - xhat * M2_g)`` with the group means
    ``M1 = mean_g(gamma * dz)``, ``M2 = mean_g(gamma * dz * xhat)`` — the GroupNorm-backward kernel evaluates
    ``gamma_c * rstd * (dz - k1 - xhat * k2)``, so it is fed ``k = M / gamma_c`` (gamma clamped away from zero, where the
    product ``gamma_c * k`` stays exact in the kernel's f64 coefficient set-up)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, slope, num_groups, eps, c_real):
        if abs(float(eps) - 1e-5) > 1e-12:
            _unsupported("GroupNorm eps != 1e-5")
        n, cp = int(x.shape[0]), int(x.shape[4])
        v = int(x.shape[1] * x.shape[2] * x.shape[3])
        g_, cpg = int(num_groups), int(c_real) // int(num_groups)
        dev = x.device
        stats = torch.zeros((n, 2, cp), device=dev, dtype=torch.float64)
        for i in range(n):
            _k_channel_stats(x[i], stats[i])
        m = float(v * cpg)
        pooled = stats[:, :, :c_real].reshape(n, 2, g_, cpg).sum(-1)                      # [n, 2, G]
        mean_g = pooled[:, 0] / m
        var_g = (pooled[:, 1] / m - mean_g * mean_g).clamp_min(0.0)
        rstd_g = 1.0 / torch.sqrt(var_g + 1e-5)
        mean_c = torch.zeros((n, cp), device=dev, dtype=torch.float64)
        rstd_c = torch.zeros((n, cp), device=dev, dtype=torch.float64)
        mean_c[:, :c_real] = mean_g.repeat_interleave(cpg, dim=1)
        rstd_c[:, :c_real] = rstd_g.repeat_interleave(cpg, dim=1)
        g64 = torch.zeros(cp, device=dev, dtype=torch.float64)
        b64 = torch.zeros(cp, device=dev, dtype=torch.float64)
        g64[:c_real], b64[:c_real] = gamma.detach().double(), beta.detach().double()
        scale = (g64 * rstd_c).float().contiguous()                                       # [n, cp]; padded channels: 0
        shift = (b64 - mean_c * g64 * rstd_c).float().contiguous()
        slope_f = slope.detach().reshape(-1)[:1].float().contiguous()
        out = torch.empty_like(x)
        for i in range(n):
            _k_bn_act_fwd(x[i], scale[i], shift[i], slope_f, out[i])
        ctx.save_for_backward(x, scale, shift, mean_c.float().contiguous(), rstd_c.float().contiguous(), slope_f, g64, mean_g, var_g)
        ctx.cfg = (n, cp, v, g_, cpg, int(c_real))
        return out

    @staticmethod
    def backward(ctx, dy):
        x, scale, shift, mean_f, rstd_f, slope_f, g64, mean_g, var_g = ctx.saved_tensors
        n, cp, v, g_, cpg, c_real = ctx.cfg
        dev = x.device
        dy = dy.contiguous()
        dz = torch.empty_like(x)
        red = torch.zeros((n, 2 * cp + 1), device=dev, dtype=torch.float64)
        for i in range(n):
            _k_bn_act_bwd(dy[i], x[i], scale[i], shift[i], mean_f[i], rstd_f[i], slope_f, dz[i], red[i])
        s1, s2 = red[:, :cp], red[:, cp:2 * cp]
        dbeta = s1[:, :c_real].sum(0).float()
        dgamma = s2[:, :c_real].sum(0).float()
        dslope = red[:, 2 * cp].sum().float().reshape(1)
        m = float(v * cpg)
        m1 = (g64 * s1)[:, :c_real].reshape(n, g_, cpg).sum(-1) / m                       # [n, G]
        m2 = (g64 * s2)[:, :c_real].reshape(n, g_, cpg).sum(-1) / m
        tiny = 1e-20
        g_safe = torch.where(g64.abs() < tiny, torch.full_like(g64, tiny), g64)
        stats_k = torch.zeros((n, 2, cp), device=dev, dtype=torch.float64)
        gstats_k = torch.zeros((n, 2, cp), device=dev, dtype=torch.float64)
        mean_c = mean_g.repeat_interleave(cpg, dim=1)
        var_c = var_g.repeat_interleave(cpg, dim=1)
        stats_k[:, 0, :c_real] = mean_c * v
        stats_k[:, 1, :c_real] = (var_c + mean_c * mean_c) * v
        gstats_k[:, 0, :c_real] = m1.repeat_interleave(cpg, dim=1) / g_safe[:c_real] * v
        gstats_k[:, 1, :c_real] = m2.repeat_interleave(cpg, dim=1) / g_safe[:c_real] * v
        gamma_k = g_safe.float().contiguous()
        gamma_k[c_real:] = 0.0                                                            # padded channels: dx stays 0
        dx = torch.empty_like(x)
        dsum = torch.zeros(cp, device=dev, dtype=torch.float64)
        _k_gn_bwd(dz, x, stats_k, gstats_k, gamma_k, dx, dsum)
        return dx, dgamma, dbeta, dslope, None, None, None


# ----------------------------------------------------------------------------- MONAI-named module tree
def _unsupported(what):
    raise NotImplementedError(f"pcb200 monai_unet: {what} is not implemented in the B200 engine yet "
                              "(3-D, norm='batch' | 'group' | 'instance', upsample_mode='deconv' | 'nontrainable').")


class ADN(nn.Sequential):
    """MONAI ``ADN`` with ordering "NDA" (children ``N``, ``D``, ``A``).  ``norm``: ``"batch"``, ``"group"`` (``GroupNormActFn``) or ``"instance"`` —
    ``InstanceNorm3d`` (no affine, no running statistics, as MONAI builds it) is BatchNorm over a batch of ONE, so each
    sample goes through the same statistics / normalise+PReLU kernels with its own statistics.  ``dropout`` > 0 is the
    identity at inference and a mask behind the fused kernel in training (see ``forward``)."""

    def __init__(self, channels: int, dropout, norm: str = "batch", num_groups: int = 8):
        super().__init__()
        self.norm_kind = norm
        if norm == "group":
            self.add_module("N", nn.GroupNorm(int(num_groups), channels))       # torch refuses channels % num_groups != 0, as MONAI does
        else:
            self.add_module("N", nn.BatchNorm3d(channels) if norm == "batch" else nn.InstanceNorm3d(channels))
        if dropout is not None:
            self.add_module("D", nn.Dropout(float(dropout)))
        self.add_module("A", nn.PReLU())
        if norm == "instance":      # constants the kernels read as gamma / beta / running statistics (not in the state_dict)
            self.register_buffer("_one", torch.ones(channels), persistent=False)
            self.register_buffer("_zero", torch.zeros(channels), persistent=False)
            self.register_buffer("_rm", torch.zeros(channels), persistent=False)
            self.register_buffer("_rv", torch.ones(channels), persistent=False)

    def forward(self, x):
        y = self._norm_act(x)
        drop = getattr(self, "D", None)
        if drop is not None and drop.p > 0 and self.training:
            # "NDA" puts the dropout between norm and PReLU; PReLU is positively homogeneous (PReLU(c z) = c PReLU(z) for
            # c >= 0) and the mask only scales by 0 or 1/(1-p), so PReLU(dropout(z)) == dropout(PReLU(z)) for the same mask:
            # the mask is applied behind the fused norm+PReLU kernel.  Padded channels are zero and stay zero.  The mask comes
            # from torch's generator on the channels-LAST tensor: same distribution as the reference's, not the same bits.
            y = torch.nn.functional.dropout(y, float(drop.p), True)
        return y

    def _norm_act(self, x):
        bn = self.N
        if self.norm_kind == "group":
            return GroupNormActFn.apply(x, bn.weight, bn.bias, self.A.weight, int(bn.num_groups), float(bn.eps), int(bn.num_channels))
        if self.norm_kind == "instance":
            outs = [BnActFn.apply(x[n:n + 1], self._one, self._zero, self.A.weight, self._rm, self._rv, True, 0.0,
                                  float(bn.eps), int(bn.num_features)) for n in range(int(x.shape[0]))]
            return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)
        return BnActFn.apply(x, bn.weight, bn.bias, self.A.weight, bn.running_mean, bn.running_var,
                             bool(self.training and bn.track_running_stats), float(bn.momentum or 0.1), float(bn.eps),
                             int(bn.num_features))


class Convolution(nn.Sequential):
    def __init__(self, in_channels, out_channels, strides=1, kernel_size=3, dropout=0.0, bias=True, conv_only=False,
                 is_transposed=False, norm: str = "batch", num_groups: int = 8):
        super().__init__()
        pad = (kernel_size - 1) // 2
        if is_transposed:
            conv = nn.ConvTranspose3d(in_channels, out_channels, kernel_size, stride=strides, padding=pad,
                                      output_padding=strides - 1, bias=bias)
        else:
            conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride=strides, padding=pad, bias=bias)
        self.add_module("conv", conv)
        if not conv_only:
            self.add_module("adn", ADN(out_channels, dropout, norm, num_groups))
        self._cfg = (kernel_size, strides, pad, bool(is_transposed))

    def forward(self, x):
        k, s, p, tr = self._cfg
        y = ConvFn.apply(x, self.conv.weight, self.conv.bias, k, s, p, tr)
        if hasattr(self, "adn"):
            if self.training and getattr(self.adn.N, "track_running_stats", False) and self.adn.norm_kind == "batch":
                with torch.no_grad():
                    self.adn.N.num_batches_tracked += 1
            y = self.adn(y)
        return y


class ResidualUnit(nn.Module):
    def __init__(self, in_channels, out_channels, strides=1, kernel_size=3, subunits=2, dropout=0.0, bias=True,
                 last_conv_only=False, norm: str = "batch", num_groups: int = 8):
        super().__init__()
        self.conv = nn.Sequential()
        self.residual: nn.Module = nn.Identity()
        sch, sst = in_channels, strides
        for su in range(max(1, subunits)):
            only = last_conv_only and su == max(1, subunits) - 1
            self.conv.add_module(f"unit{su:d}", Convolution(sch, out_channels, sst, kernel_size, dropout, bias, only, norm=norm,
                                                                num_groups=num_groups))
            sch, sst = out_channels, 1
        self._res_cfg = None
        if strides != 1 or in_channels != out_channels:
            rk, rp = (kernel_size, (kernel_size - 1) // 2) if strides != 1 else (1, 0)
            self.residual = nn.Conv3d(in_channels, out_channels, rk, strides, rp, bias=bias)
            self._res_cfg = (rk, strides, rp)

    def forward(self, x):
        cx = self.conv(x)
        if self._res_cfg is None:
            res = x
        else:
            rk, rs, rp = self._res_cfg
            res = ConvFn.apply(x, self.residual.weight, self.residual.bias, rk, rs, rp, False)
        return cx + res


class SkipConnection(nn.Module):
    def __init__(self, submodule: nn.Module, c_in: int):
        super().__init__()
        self.submodule = submodule
        self._c_in = c_in

    def forward(self, x):
        y = self.submodule(x)
        # cat([x, y], channel) on the REAL channels, re-padded to a multiple of 16 for the consuming conv
        cx, cy = self._c_in, self._c_out
        cat = torch.cat([x[..., :cx], y[..., :cy]], dim=-1)
        padc = _pad16(cx + cy) - (cx + cy)
        if padc:
            cat = torch.nn.functional.pad(cat, (0, padc))
        return cat.contiguous()


class UpSample(nn.Sequential):
    """monai.networks.blocks.UpSample, ``mode="nontrainable"`` (what the reference's ``UpsampleModeUNet`` swaps in for the
    transposed conv, ``monai_models.py:104-139``): children ``preconv`` (Conv3d kernel 1, only when the channel counts differ)
    and ``upsample_non_trainable`` (``nn.Upsample``; the linear family means trilinear in 3-D).  The 1x1 conv runs on the
    implicit-GEMM kernel; the interpolation itself is a device tensor op on the padded channels-last tensor (no hand-written
    resampling kernel yet) — padded channels are zero and stay zero."""

    def __init__(self, in_channels: int, out_channels: int, scale_factor: int, interp_mode: str = "linear",
                 align_corners: bool = True, bias: bool = True):
        super().__init__()
        if out_channels != in_channels:
            self.add_module("preconv", nn.Conv3d(in_channels, out_channels, kernel_size=1, bias=bias))
        interp = str(interp_mode).lower()
        if interp in ("linear", "bilinear", "trilinear"):
            interp = "trilinear"
        elif interp != "nearest":
            _unsupported(f"upsample_interp_mode={interp_mode!r}")
        self.add_module("upsample_non_trainable", nn.Upsample(scale_factor=(float(scale_factor),) * 3, mode=interp,
                                                              align_corners=align_corners))

    def forward(self, x):
        if hasattr(self, "preconv"):
            x = ConvFn.apply(x, self.preconv.weight, self.preconv.bias, 1, 1, 0, False)
        up = self.upsample_non_trainable
        y = torch.nn.functional.interpolate(x.permute(0, 4, 1, 2, 3), scale_factor=up.scale_factor, mode=up.mode,
                                            align_corners=up.align_corners)
        return y.permute(0, 2, 3, 4, 1).contiguous()


class UNet(nn.Module):
    """monai.networks.nets.UNet (3-D, PReLU; batch / instance / group norm) on the B200 engine, with the reference's
    ``UpsampleModeUNet`` switch (``monai_models.py:84-139``): ``upsample_mode="deconv"`` (MONAI's transposed conv + ADN) or
    ``"nontrainable"`` (``UpSample``: 1x1 conv + interpolation, no norm / activation)."""

    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, channels: Sequence[int], strides: Sequence[int],
                 kernel_size: int = 3, up_kernel_size: int = 3, num_res_units: int = 0, norm="batch", dropout: float = 0.0,
                 bias: bool = True, upsample_mode: str = "deconv", upsample_interp_mode: str = "linear",
                 upsample_align_corners: bool = True):
        super().__init__()
        self.upsample_mode = str(upsample_mode or "deconv").lower()
        if self.upsample_mode not in ("deconv", "nontrainable"):
            _unsupported(f"upsample_mode={upsample_mode!r}")
        self.upsample_interp_mode, self.upsample_align_corners = upsample_interp_mode, upsample_align_corners
        if spatial_dims != 3:
            _unsupported("spatial_dims != 3")
        norm_kw = dict(norm[1]) if isinstance(norm, (tuple, list)) and len(norm) > 1 else {}
        norm = str(norm[0] if isinstance(norm, (tuple, list)) else norm).lower()
        if norm not in ("batch", "instance", "group"):
            _unsupported(f"norm={norm!r}")
        if norm == "group" and "num_groups" not in norm_kw:
            raise TypeError("GroupNorm.__init__() missing 1 required positional argument: 'num_groups'")   # MONAI / torch behaviour
        self.norm, self.num_groups = norm, int(norm_kw.get("num_groups", 8))
        if kernel_size != 3 or up_kernel_size != 3:
            _unsupported("kernel_size != 3")
        if len(channels) < 2:
            raise ValueError("the length of `channels` should be no less than 2.")
        if len(strides) < len(channels) - 1:
            raise ValueError("the length of `strides` should equal to `len(channels) - 1`.")
        if any(s not in (1, 2) for s in strides):
            _unsupported("strides other than 1 or 2")
        self.dimensions, self.kernel_size, self.num_res_units = spatial_dims, kernel_size, num_res_units
        self.in_channels, self.out_channels, self.dropout, self.bias = in_channels, out_channels, dropout, bias
        self.n_down = sum(1 for s in strides[: len(channels) - 1] if s == 2)

        def block(inc, outc, ch, st, is_top):
            c, s = ch[0], st[0]
            if len(ch) > 2:
                sub, upc, sub_out = block(c, c, ch[1:], st[1:], False), c * 2, c
            else:
                sub, upc, sub_out = self._down(c, ch[1], 1), c + ch[1], ch[1]
            skip = SkipConnection(sub, c)
            skip._c_out = sub_out
            return nn.Sequential(self._down(inc, c, s), skip, self._up(upc, outc, s, is_top))

        self.model = block(in_channels, out_channels, list(channels), list(strides), True)

    def _down(self, i, o, s):
        if self.num_res_units > 0:
            return ResidualUnit(i, o, s, self.kernel_size, self.num_res_units, self.dropout, self.bias, norm=self.norm,
                                num_groups=self.num_groups)
        return Convolution(i, o, s, self.kernel_size, self.dropout, self.bias, norm=self.norm, num_groups=self.num_groups)

    def _up(self, i, o, s, is_top):
        if self.upsample_mode == "nontrainable":
            conv = UpSample(i, o, s, self.upsample_interp_mode, self.upsample_align_corners, self.bias)
        else:
            conv = Convolution(i, o, s, 3, self.dropout, self.bias, conv_only=is_top and self.num_res_units == 0,
                               is_transposed=True, norm=self.norm, num_groups=self.num_groups)
        if self.num_res_units > 0:
            return nn.Sequential(conv, ResidualUnit(o, o, 1, self.kernel_size, 1, self.dropout, self.bias, last_conv_only=is_top,
                                                    norm=self.norm, num_groups=self.num_groups))
        return conv

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        L.require_device(x, "monai_unet forward")
        if x.dim() != 5:
            raise ValueError(f"monai_unet expects (B, C, D, H, W); got shape {tuple(x.shape)}")
        div = 2 ** self.n_down
        if any(int(s) % div for s in x.shape[2:]):
            raise ValueError(f"monai_unet input spatial size must be divisible by {div}, got {tuple(x.shape[2:])}")
        c = int(x.shape[1])
        h = x.permute(0, 2, 3, 4, 1)
        padc = _pad16(c) - c
        if padc:
            h = torch.nn.functional.pad(h, (0, padc))
        h = h.to(_BF16).contiguous()
        y = self.model(h)
        return y[..., : self.out_channels].permute(0, 4, 1, 2, 3).to(x.dtype if x.dtype != torch.float64 else torch.float32).contiguous()


class MONAIModelWrapper(ConnectomicsModel):
    """``monai_models.py:29-56`` — ConnectomicsModel interface.  The reference's wrapper also squeezes a singleton depth in
    front of its 2-D nets and restores it behind them; this engine builds 3-D nets only (``spatial_dims != 3`` is refused at
    construction), which take and return ``(B, C, D, H, W)`` as they are."""

    def __init__(self, model: nn.Module):
        super().__init__()
        self.model = model
        self.supports_deep_supervision = False
        self.output_scales = 1

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.model(x)


@register_architecture("monai_unet")
def build_monai_unet(cfg) -> ConnectomicsModel:
    """MONAI UNet with residual units on the B200 engine (``monai_models.py:197-250``)."""
    m = cfg.model.monai
    size = getattr(cfg.model, "input_size", None)
    dims = len(size) if size else getattr(m, "spatial_dims", 3)
    channels = list(getattr(m, "filters", [32, 64, 128, 256, 512]))
    norm = getattr(m, "norm", "batch")
    if norm == "group":                                   # monai_models.py:74-81 _resolve_norm
        norm = ("group", {"num_groups": getattr(m, "num_groups", 8)})
    model = UNet(spatial_dims=dims, in_channels=cfg.model.in_channels, out_channels=cfg.model.out_channels,
                 channels=channels, strides=[2] * (len(channels) - 1), num_res_units=getattr(m, "num_res_units", 2),
                 kernel_size=getattr(m, "kernel_size", 3), norm=norm, dropout=getattr(m, "dropout", 0.0),
                 upsample_mode=getattr(m, "upsample_mode", "deconv"),
                 upsample_interp_mode=getattr(m, "upsample_interp_mode", "linear"),
                 upsample_align_corners=getattr(m, "upsample_align_corners", True))
    return MONAIModelWrapper(model)


__all__ = ["UNet", "MONAIModelWrapper", "build_monai_unet"]
