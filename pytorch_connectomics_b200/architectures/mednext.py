"""MedNeXt on the B200 engine — drop-in for what
``connectomics/models/architectures/mednext_models.py`` builds from the third-party
``nnunet_mednext`` package (``:23-32,374-380,479``).

Module / attribute names and ``state_dict`` keys are the upstream ones
(``stem``, ``enc_block_{l}.{i}.{conv1,norm,conv2,conv3}``, ``down_{l}``/``up_{l}`` with ``res_conv``,
``bottleneck``, ``dec_block_{l}``, ``out_{k}.conv_out``, ``dummy_tensor``), so reference checkpoints
load unchanged (``training/model_weights.py:74`` loads into ``model.model``).  The ``nn.Conv3d`` /
``nn.GroupNorm`` children are *parameter containers only* (same shapes, same default init as
upstream); the arithmetic runs in hand-written sm_100a kernels through the C ABI
(``include/pcb200.h``): activations travel channels-last bf16 ``[N,D,H,W,C]`` between kernels.

There is no PyTorch/CPU fallback: ``forward`` on a non-CUDA tensor raises ``RuntimeError``.
"""

from __future__ import annotations

import ctypes
from typing import Any, Dict, List, Mapping, Optional, Sequence, Union

import torch
import torch.nn as nn

from .. import _lib as L
from .base import ConnectomicsModel
from .registry import register_architecture
from . import _mednext_ops as ops


def _unsupported(what: str):
    raise NotImplementedError(
        f"pcb200 MedNeXt: {what} is not implemented in the B200 engine yet (3-D / 2-D, GroupNorm(C groups) | LayerNorm, optional GRN).")


# ----------------------------------------------------------------------------- dim="2d" (upstream blocks.py: Conv2d / ConvTranspose2d)
# A 2-D network is the 3-D one on a volume of depth 1.  With zero padding only the centre z-plane of a k^3 stencil ever meets
# data (same, stride-2 and transposed stride-2 alike: out z = 0 pairs input z = 0 with tap k//2), so the 2-D depthwise weight
# [C,1,k,k] is lifted to [C,1,k,k,k] with zero off-centre planes and every kernel runs unchanged; GroupNorm statistics, the
# 1x1 convolutions and the heads see the same H*W voxels.  The lift is a differentiable view (unsqueeze + pad), so autograd
# returns the centre plane of the 3-D weight gradient to the 2-D parameter.  The one geometric difference is the up block:
# the 3-D kernel front-pads its output along z as well (depth 2, plane 0 = padding) — plane 1 is the 2-D result.
def _lift_dw(w: torch.Tensor) -> torch.Tensor:
    p = int(w.shape[-1]) // 2
    return torch.nn.functional.pad(w.unsqueeze(2), (0, 0, 0, 0, p, p))


def _conv(dim: str):
    return nn.Conv2d if dim == "2d" else nn.Conv3d


def _convt(dim: str):
    return nn.ConvTranspose2d if dim == "2d" else nn.ConvTranspose3d


class LayerNorm(nn.Module):
    """upstream blocks.py::LayerNorm(data_format="channels_first"), eps 1e-5 — parameter container (``weight``,
    ``bias``; same ``state_dict`` keys as the GroupNorm variant); the arithmetic is ``csrc/layernorm.cu``."""

    def __init__(self, normalized_shape: int, eps: float = 1e-5, data_format: str = "channels_first"):
        super().__init__()
        if abs(eps - 1e-5) > 1e-12:
            _unsupported("LayerNorm eps != 1e-5")
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        self.eps = eps
        self.data_format = data_format
        self.normalized_shape = (normalized_shape,)


class MedNeXtBlock(nn.Module):
    """upstream blocks.py::MedNeXtBlock — conv1 (depthwise k^3) -> GroupNorm(C groups) -> conv2 (1x1,
    C->rC) -> GELU -> conv3 (1x1, rC->Cout) [+ x].  ``forward`` takes/returns channels-last bf16."""

    _dw_mode = L.DW_SAME

    def __init__(self, in_channels: int, out_channels: int, exp_r: int = 4, kernel_size: int = 7,
                 do_res: bool = True, norm_type: str = "group", n_groups=None, dim: str = "3d",
                 grn: bool = False):
        super().__init__()
        if dim not in ("2d", "3d"):
            raise ValueError(f"dim must be '2d' or '3d', got {dim!r}")
        if norm_type not in ("group", "layer"):
            raise ValueError(f"norm_type must be 'group' or 'layer', got {norm_type!r}")
        if n_groups is not None and n_groups != in_channels:
            _unsupported("n_groups != in_channels")
        self.do_res = do_res
        self.dim = dim
        self.grn = grn
        conv = _conv(dim)
        self.conv1 = conv(in_channels, in_channels, kernel_size, 1, kernel_size // 2, groups=in_channels)
        self.norm_type = norm_type
        self.norm = (nn.GroupNorm(num_groups=in_channels, num_channels=in_channels) if norm_type == "group"
                     else LayerNorm(in_channels))
        self.conv2 = conv(in_channels, exp_r * in_channels, 1)
        self.act = nn.GELU()
        self.conv3 = conv(exp_r * in_channels, out_channels, 1)
        if grn:      # upstream blocks.py: zero-initialised (1, rC, 1, 1[, 1]) parameters
            shape = (1, exp_r * in_channels) + (1,) * (3 if dim == "3d" else 2)
            self.grn_beta = nn.Parameter(torch.zeros(shape), requires_grad=True)
            self.grn_gamma = nn.Parameter(torch.zeros(shape), requires_grad=True)

    def _params(self) -> List[torch.Tensor]:
        w1 = _lift_dw(self.conv1.weight) if self.dim == "2d" else self.conv1.weight     # 1x1 weights reshape to [O, I] as they are
        p = [w1, self.conv1.bias, self.norm.weight, self.norm.bias,
             self.conv2.weight, self.conv2.bias, self.conv3.weight, self.conv3.bias]
        rc = getattr(self, "res_conv", None)
        if rc is not None:
            p += [rc.weight, rc.bias]
        return p

    def forward(self, x: torch.Tensor, skip: Optional[torch.Tensor] = None) -> torch.Tensor:
        has_rc = getattr(self, "res_conv", None) is not None
        plane = self.dim == "2d" and self._dw_mode == L.DW_UP
        if plane and skip is not None:      # depth-2 skip whose plane 1 is the encoder feature (plane 0 only meets padding)
            skip = ops.as_channels_last_2d(skip)
            skip = ops._mark(torch.cat([torch.zeros_like(skip), skip], dim=1))
        if self.grn:
            out = self._forward_grn(x, skip, has_rc)
        else:
            out = ops.block_apply(x, skip, self._params(), self._dw_mode, self.conv1.kernel_size[0],
                                  bool(self.do_res), has_rc, self.norm_type)
        return ops._mark(out[:, 1:2].contiguous()) if plane else out

    def _forward_grn(self, x: torch.Tensor, skip: Optional[torch.Tensor], has_rc: bool) -> torch.Tensor:
        """Block with Global Response Normalisation (upstream blocks.py: between GELU and conv3,
        ``gx = ||h||_2 over space; nx = gx / (mean_channels gx + 1e-6); h <- gamma * (h * nx) + beta + h``).

        The global reduction sits in the middle of what the fused block kernel keeps on chip, so a GRN block is COMPOSED
        kernel by kernel: stencil (``pcb_dwconv_fwd``) -> norm kernels -> conv2 (``pcb_pw_fwd``, tcgen05) -> GELU -> GRN
        statistics -> conv3 (``pcb_pw_fwd``) -> residual.  GRN itself never touches the expanded tensor again: per sample it is
        a per-channel scale ``s = gamma * nx + 1`` and a shift ``beta``, folded into conv3 as ``W3 diag(s)`` and
        ``b3 + W3 beta``.  GELU, the norm over space and the residual adds are elementwise / reduction tensor ops on the
        device (not hand-written kernels yet), the expanded activation makes one HBM round trip, and backward is autograd over
        the kernels' own backward functions — this path is for coverage of the option, not the measured one."""
        F = torch.nn.functional
        mode, k = self._dw_mode, int(self.conv1.kernel_size[0])
        p = self._params()
        x = ops.as_channels_last(x)
        y = ops.dwconv_apply(x, p[0], p[1], mode, k)
        a = ops.norm_apply(y, p[2], p[3], self.norm_type)
        h = ops._mark(F.gelu(ops.pointwise_apply(a, p[4], p[5])))
        hid, co = int(p[4].shape[0]), int(p[6].shape[0])
        gx = torch.linalg.vector_norm(h, ord=2, dim=(1, 2, 3), dtype=torch.float32)              # [N, rC]
        nx = gx / (gx.mean(dim=1, keepdim=True) + 1e-6)
        s = self.grn_gamma.reshape(1, hid).float() * nx + 1.0
        w3 = p[6].reshape(co, hid).float()
        b3 = p[7].float() + w3 @ self.grn_beta.reshape(hid).float()
        outs = [ops.pointwise_apply(ops._mark(h[n:n + 1]), w3 * s[n][None, :], b3) for n in range(int(h.shape[0]))]
        o = outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)
        if mode == L.DW_SAME:
            if self.do_res:
                o = o + x
            return ops._mark(o)
        if has_rc:
            wr, br = self.res_conv.weight, self.res_conv.bias
            if mode == L.DW_DOWN:       # Conv(C, Co, 1, stride 2): the even-coordinate voxels through the 1x1 GEMM
                o = o + ops.pointwise_apply(ops._mark(x[:, ::2, ::2, ::2].contiguous()), wr, br)
            else:                       # ConvTranspose(C, Co, 1, stride 2): even output voxels get W^T x + b, the others b
                even = ops.pointwise_apply(x, wr.reshape(wr.shape[0], wr.shape[1]).t(), br)
                r = br.to(o.dtype).reshape(1, 1, 1, 1, co).expand(o.shape).clone()
                r[:, ::2, ::2, ::2] = even
                o = o + r
        if mode == L.DW_UP:
            o = F.pad(o, (0, 0, 1, 0, 1, 0, 1, 0))      # one voxel at the FRONT of every spatial axis
            if skip is not None:
                skip = ops.as_channels_last(skip)
                if tuple(skip.shape) != tuple(o.shape):
                    raise ValueError(f"skip shape {tuple(skip.shape)} != up-block output shape {tuple(o.shape)}")
                o = o + skip
        return ops._mark(o)


class MedNeXtDownBlock(MedNeXtBlock):
    """upstream blocks.py::MedNeXtDownBlock — stride-2 depthwise conv1; residual = Conv3d(k=1, stride 2)."""

    _dw_mode = L.DW_DOWN

    def __init__(self, in_channels, out_channels, exp_r=4, kernel_size=7, do_res=False,
                 norm_type="group", dim="3d", grn=False):
        super().__init__(in_channels, out_channels, exp_r, kernel_size, do_res=False,
                         norm_type=norm_type, dim=dim, grn=grn)
        self.resample_do_res = do_res
        if do_res:
            self.res_conv = _conv(dim)(in_channels, out_channels, 1, stride=2)
        self.conv1 = _conv(dim)(in_channels, in_channels, kernel_size, 2, kernel_size // 2, groups=in_channels)


class MedNeXtUpBlock(MedNeXtBlock):
    """upstream blocks.py::MedNeXtUpBlock — transposed stride-2 depthwise conv1 (spatial 2s-1), block
    output zero-padded by one voxel at the front of each axis; residual = ConvTranspose3d(k=1, stride 2)."""

    _dw_mode = L.DW_UP

    def __init__(self, in_channels, out_channels, exp_r=4, kernel_size=7, do_res=False,
                 norm_type="group", dim="3d", grn=False):
        super().__init__(in_channels, out_channels, exp_r, kernel_size, do_res=False,
                         norm_type=norm_type, dim=dim, grn=grn)
        self.resample_do_res = do_res
        if do_res:
            self.res_conv = _convt(dim)(in_channels, out_channels, 1, stride=2)
        self.conv1 = _convt(dim)(in_channels, in_channels, kernel_size, 2, kernel_size // 2, groups=in_channels)


class OutBlock(nn.Module):
    """upstream blocks.py::OutBlock — ConvTranspose3d(C, n_classes, k=1) (ConvTranspose2d for ``dim="2d"``).  channels-last
    bf16 in, NCDHW out (NCHW for 2-D: the depth-1 axis is dropped)."""

    def __init__(self, in_channels, n_classes, dim="3d"):
        super().__init__()
        if dim not in ("2d", "3d"):
            raise ValueError(f"dim must be '2d' or '3d', got {dim!r}")
        self.dim = dim
        self.conv_out = _convt(dim)(in_channels, n_classes, 1)

    def forward(self, x: torch.Tensor, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
        if self.dim == "2d":
            x = ops.as_channels_last_2d(x)
        y = ops.head_apply(x, self.conv_out.weight, self.conv_out.bias, out_dtype)
        return y.squeeze(2) if self.dim == "2d" else y


class MedNeXt(nn.Module):
    """upstream MedNextV1.py::MedNeXt (+ the fork's forward_features / forward_output)."""

    def __init__(self, in_channels: int, n_channels: int, n_classes: int,
                 exp_r: Union[int, Sequence[int]] = 4, kernel_size: int = 7,
                 enc_kernel_size: int = None, dec_kernel_size: int = None,
                 deep_supervision: bool = False, do_res: bool = False, do_res_up_down: bool = False,
                 checkpoint_style: str = None, block_counts: Sequence[int] = (2,) * 9,
                 norm_type: str = "group", dim: str = "3d", grn: bool = False):
        super().__init__()
        self.do_ds = deep_supervision
        if checkpoint_style not in (None, "outside_block"):
            raise ValueError(f"checkpoint_style must be None or 'outside_block', got {checkpoint_style!r}")
        # The engine always recomputes the expanded tensor in backward and stores only the block input
        # and the depthwise output, so the flag is kept for API compatibility and changes nothing.
        self.inside_block_checkpointing = False
        self.outside_block_checkpointing = checkpoint_style == "outside_block"
        if dim not in ("2d", "3d"):
            raise ValueError(f"dim must be '2d' or '3d', got {dim!r}")
        self.dim = dim
        if kernel_size is not None:
            enc_kernel_size = dec_kernel_size = kernel_size
        if n_channels % 16 != 0:
            raise ValueError(f"pcb200 MedNeXt needs base_channels to be a multiple of 16, got {n_channels}")
        exp_r = [exp_r] * len(block_counts) if isinstance(exp_r, int) else list(exp_r)
        n = n_channels
        kw = dict(norm_type=norm_type, dim=dim, grn=grn)
        self.stem = _conv(dim)(in_channels, n, 1)

        def stage(c, i, k):
            return nn.Sequential(*[MedNeXtBlock(c, c, exp_r[i], k, do_res=do_res, **kw)
                                   for _ in range(block_counts[i])])

        ek, dk = enc_kernel_size, dec_kernel_size
        self.enc_block_0 = stage(n, 0, ek)
        self.down_0 = MedNeXtDownBlock(n, 2 * n, exp_r[1], ek, do_res=do_res_up_down, **kw)
        self.enc_block_1 = stage(2 * n, 1, ek)
        self.down_1 = MedNeXtDownBlock(2 * n, 4 * n, exp_r[2], ek, do_res=do_res_up_down, **kw)
        self.enc_block_2 = stage(4 * n, 2, ek)
        self.down_2 = MedNeXtDownBlock(4 * n, 8 * n, exp_r[3], ek, do_res=do_res_up_down, **kw)
        self.enc_block_3 = stage(8 * n, 3, ek)
        self.down_3 = MedNeXtDownBlock(8 * n, 16 * n, exp_r[4], ek, do_res=do_res_up_down, **kw)
        self.bottleneck = stage(16 * n, 4, dk)
        self.up_3 = MedNeXtUpBlock(16 * n, 8 * n, exp_r[5], dk, do_res=do_res_up_down, **kw)
        self.dec_block_3 = stage(8 * n, 5, dk)
        self.up_2 = MedNeXtUpBlock(8 * n, 4 * n, exp_r[6], dk, do_res=do_res_up_down, **kw)
        self.dec_block_2 = stage(4 * n, 6, dk)
        self.up_1 = MedNeXtUpBlock(4 * n, 2 * n, exp_r[7], dk, do_res=do_res_up_down, **kw)
        self.dec_block_1 = stage(2 * n, 7, dk)
        self.up_0 = MedNeXtUpBlock(2 * n, n, exp_r[8], dk, do_res=do_res_up_down, **kw)
        self.dec_block_0 = stage(n, 8, dk)
        self.out_0 = OutBlock(n, n_classes, dim)
        self.dummy_tensor = nn.Parameter(torch.tensor([1.0]), requires_grad=True)
        if deep_supervision:
            self.out_1 = OutBlock(2 * n, n_classes, dim)
            self.out_2 = OutBlock(4 * n, n_classes, dim)
            self.out_3 = OutBlock(8 * n, n_classes, dim)
            self.out_4 = OutBlock(16 * n, n_classes, dim)
        self.block_counts = list(block_counts)
        self.output_dtype: Optional[torch.dtype] = None   # None -> same dtype as the input

    # ---- channels-last trunk
    def _trunk(self, x: torch.Tensor) -> List[torch.Tensor]:
        L.require_device(x, "MedNeXt.forward")
        if self.dim == "2d":
            if x.dim() != 4:
                raise ValueError(f"MedNeXt (dim='2d') expects (B, C, H, W); got shape {tuple(x.shape)}")
        elif x.dim() != 5:
            raise ValueError(f"MedNeXt expects (B, C, D, H, W); got shape {tuple(x.shape)}")
        if any(int(s) % 16 for s in x.shape[2:]):
            raise ValueError(f"MedNeXt input spatial size must be divisible by 16, got {tuple(x.shape[2:])}")
        if self.dim == "2d":
            x = x.unsqueeze(2)        # depth-1 volume; the 1x1 stem weight [n, Cin, 1, 1] flattens to the same [n, Cin]
        x = ops.stem_apply(x, self.stem.weight, self.stem.bias)
        r0 = self.enc_block_0(x)
        x = self.down_0(r0)
        r1 = self.enc_block_1(x)
        x = self.down_1(r1)
        r2 = self.enc_block_2(x)
        x = self.down_2(r2)
        r3 = self.enc_block_3(x)
        x = self.down_3(r3)
        b = self.bottleneck(x)
        d3 = self.dec_block_3(self.up_3(b, r3))       # skip add fused into the up block's epilogue
        d2 = self.dec_block_2(self.up_2(d3, r2))
        d1 = self.dec_block_1(self.up_1(d2, r1))
        f0 = self.dec_block_0(self.up_0(d1, r0))
        return [f0, d1, d2, d3, b]

    def _odt(self, x: torch.Tensor) -> torch.dtype:
        return self.output_dtype or (x.dtype if x.dtype in (torch.float16, torch.bfloat16, torch.float32)
                                     else torch.float32)

    def forward_features(self, x: torch.Tensor) -> torch.Tensor:
        """Shared full-resolution feature map [B, n, D, H, W] (NCDHW view of the channels-last tensor; [B, n, H, W] for 2-D)."""
        f = self._trunk(x)[0].permute(0, 4, 1, 2, 3)
        return f.squeeze(2) if self.dim == "2d" else f

    def forward_output(self, features: torch.Tensor) -> torch.Tensor:
        cl = ops.as_channels_last_2d if self.dim == "2d" else ops.as_channels_last
        return self.out_0(cl(features), torch.float32 if features.dtype == torch.bfloat16
                          and self.output_dtype is None else (self.output_dtype or features.dtype))

    def forward(self, x: torch.Tensor):
        odt = self._odt(x)
        if self._use_native(x):
            outs = self.native_plan().forward(x, odt)
            return outs if self.do_ds else outs[0]
        f = self._trunk(x)
        y = self.out_0(f[0], odt)
        if self.do_ds:
            return [y, self.out_1(f[1], odt), self.out_2(f[2], odt), self.out_3(f[3], odt), self.out_4(f[4], odt)]
        return y

    # ---- native inference plan (include/pcb200.h pcb_net_*): no autograd, whole network enqueued by the library
    native_inference: bool = True

    def native_plan(self):
        """The ``pcb_net`` of this module (built lazily, rebuilt when a weight changed); ``None`` when the architecture
        is not served by the native plan (LayerNorm / GRN blocks)."""
        from .native import NativeMedNeXt, native_eligible
        plan = self.__dict__.get("_native_plan")
        if plan is not None and not plan.stale() and plan.device == self.stem.weight.device:
            return plan
        self.__dict__["_native_plan"] = None
        if native_eligible(self) is not None or not self.stem.weight.is_cuda:
            return None
        plan = NativeMedNeXt(self)
        self.__dict__["_native_plan"] = plan
        return plan

    def _use_native(self, x: torch.Tensor) -> bool:
        if not self.native_inference or self.training or not x.is_cuda or x.dim() != 5:
            return False
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return False              # someone may call backward: keep the autograd path
        if int(x.shape[0]) > 8 or int(x.shape[1]) != self.stem.in_channels or any(int(s) % 16 for s in x.shape[2:]):
            return False
        return self.native_plan() is not None


_V1 = {
    "S": dict(exp_r=2, block_counts=[2] * 9, checkpoint_style=None),
    "B": dict(exp_r=[2, 3, 4, 4, 4, 4, 4, 3, 2], block_counts=[2] * 9, checkpoint_style=None),
    "M": dict(exp_r=[2, 3, 4, 4, 4, 4, 4, 3, 2], block_counts=[3, 4, 4, 4, 4, 4, 4, 4, 3],
              checkpoint_style="outside_block"),
    "L": dict(exp_r=[3, 4, 8, 8, 8, 8, 8, 4, 3], block_counts=[3, 4, 8, 8, 8, 8, 8, 4, 3],
              checkpoint_style="outside_block"),
}


def create_mednext_v1(num_input_channels, num_classes, model_id, kernel_size=3, deep_supervision=False):
    """upstream create_mednext_v1.py size table (n_channels 32, do_res, do_res_up_down)."""
    s = _V1[model_id]
    return MedNeXt(num_input_channels, 32, num_classes, exp_r=s["exp_r"], kernel_size=kernel_size,
                   deep_supervision=deep_supervision, do_res=True, do_res_up_down=True,
                   block_counts=s["block_counts"], checkpoint_style=s["checkpoint_style"])


# ----------------------------------------------------------------------------- reference wrappers
class MedNeXtWrapper(ConnectomicsModel):
    """``mednext_models.py:38-89`` — list -> {"output", "ds_1".."ds_4"} when deep supervision is on."""

    def __init__(self, model: nn.Module, deep_supervision: bool = False):
        super().__init__()
        self.model = model
        self.supports_deep_supervision = deep_supervision
        self.output_scales = 5 if deep_supervision else 1

    def forward(self, x):
        out = self.model(x)
        if self.supports_deep_supervision and isinstance(out, list):
            return {"output": out[0], "ds_1": out[1], "ds_2": out[2], "ds_3": out[3], "ds_4": out[4]}
        return out

    def native_plan(self):
        """``pcb_net`` of the trunk when this wrapper returns a plain tensor (the sliding-window engine then runs its
        whole tile loop in the library, ``pcb_sw_run``); ``None`` otherwise."""
        if self.supports_deep_supervision or self.training or not getattr(self.model, "native_inference", False):
            return None
        return self.model.native_plan() if hasattr(self.model, "native_plan") else None


def _cfg_value(cfg: Any, key: str, default: Any = None) -> Any:
    return cfg.get(key, default) if isinstance(cfg, Mapping) else getattr(cfg, key, default)


class MedNeXtTaskHead(nn.Module):
    """``mednext_models.py:129-194`` — 1x1 in-projection -> N MedNeXt blocks -> 1x1 projection, on the
    shared channels-last feature map."""

    def __init__(self, in_channels: int, out_channels: int, num_blocks: int,
                 hidden_channels: Optional[int] = None, *, exp_r: int, kernel_size: int, do_res: bool,
                 norm_type: str, dim: str, grn: bool):
        super().__init__()
        if num_blocks < 0:
            raise ValueError(f"MedNeXt task head num_blocks must be >= 0, got {num_blocks}")
        if out_channels <= 0:
            raise ValueError(f"MedNeXt task head out_channels must be positive, got {out_channels}")
        hidden_channels = in_channels if hidden_channels is None else hidden_channels
        if hidden_channels <= 0:
            raise ValueError(f"MedNeXt task head hidden_channels must be positive, got {hidden_channels}")
        if hidden_channels > in_channels:
            raise ValueError("MedNeXt task head hidden_channels must not exceed the shared feature width "
                             f"({hidden_channels} > {in_channels})")
        if dim not in ("2d", "3d"):
            raise ValueError(f"MedNeXt task head dim must be '2d' or '3d', got {dim}")
        self.dim = dim
        self.input_projection = (_conv(dim)(in_channels, hidden_channels, 1)
                                 if hidden_channels != in_channels else nn.Identity())
        blocks = [MedNeXtBlock(hidden_channels, hidden_channels, exp_r, kernel_size, do_res=do_res,
                               norm_type=norm_type, dim=dim, grn=grn) for _ in range(num_blocks)]
        self.blocks = nn.Sequential(*blocks) if blocks else nn.Identity()
        self.projection = _conv(dim)(hidden_channels, out_channels, 1)
        self.hidden_channels = hidden_channels

    def forward(self, x: torch.Tensor, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
        x = ops.as_channels_last_2d(x) if self.dim == "2d" else ops.as_channels_last(x)
        if not isinstance(self.input_projection, nn.Identity):
            x = ops.pointwise_apply(x, self.input_projection.weight, self.input_projection.bias)
        x = self.blocks(x)
        # Conv3d weight is [out, in, 1,1,1]; the head kernel takes the ConvTranspose layout [in, out]
        y = ops.head_apply(x, self.projection.weight, self.projection.bias, out_dtype, conv_layout=True)
        return y.squeeze(2) if self.dim == "2d" else y


def _infer_head_block_kwargs(model: nn.Module) -> Dict[str, Any]:
    # mednext_models.py:99-126
    if not hasattr(model, "dec_block_0") or len(model.dec_block_0) == 0:
        raise ValueError("MedNeXt trunk must expose a non-empty dec_block_0 to build task heads.")
    ref = model.dec_block_0[0]
    if not isinstance(ref, MedNeXtBlock):
        raise TypeError("Expected MedNeXt dec_block_0 to contain MedNeXtBlock instances for multi-head reuse.")
    k = ref.conv1.kernel_size
    k = k[0] if isinstance(k, tuple) else k
    return {"exp_r": ref.conv2.out_channels // ref.conv2.in_channels, "kernel_size": int(k),
            "do_res": ref.do_res, "norm_type": "group" if isinstance(ref.norm, nn.GroupNorm) else "layer",
            "dim": ref.dim, "grn": ref.grn}


class MedNeXtMultiHeadWrapper(ConnectomicsModel):
    """``mednext_models.py:197-273`` — named task heads on the shared features; returns
    ``{"output": {head: tensor}}``; deep-supervision trunks are rejected."""

    def __init__(self, model: nn.Module, heads: Mapping[str, Any], *, primary_head: Optional[str] = None):
        super().__init__()
        if getattr(model, "do_ds", False):
            raise ValueError("MedNeXtMultiHeadWrapper does not support deep supervision yet. "
                             "Disable deep supervision for the trunk first.")
        if not hasattr(model, "forward_features"):
            raise ValueError("MedNeXt trunk must expose forward_features() before using MedNeXtMultiHeadWrapper.")
        if not heads:
            raise ValueError("MedNeXtMultiHeadWrapper requires at least one named task head.")
        self.model = model
        self.supports_deep_supervision = False
        self.output_scales = 1
        self.feature_channels = int(model.stem.out_channels)
        self.head_block_kwargs = _infer_head_block_kwargs(model)
        task_heads, specs = {}, {}
        for name, hc in heads.items():
            oc = int(_cfg_value(hc, "out_channels", hc))
            nb = int(_cfg_value(hc, "num_blocks", 0))
            hid = _cfg_value(hc, "hidden_channels", None)
            hid = int(hid) if hid is not None else None
            task_heads[name] = MedNeXtTaskHead(self.feature_channels, oc, nb, hid, **self.head_block_kwargs)
            specs[name] = {"out_channels": oc, "num_blocks": nb, "hidden_channels": hid or self.feature_channels}
        self.heads = nn.ModuleDict(task_heads)
        self.head_specs = specs
        primary = primary_head or next(iter(self.heads.keys()))
        if primary not in self.heads:
            raise ValueError(f"primary_head '{primary}' is not one of the configured heads: {sorted(self.heads.keys())}")
        self.primary_head = primary

    def forward_features(self, x):
        return self.model.forward_features(x)

    def forward_heads(self, features):
        odt = self.model.output_dtype or (torch.float32 if features.dtype == torch.bfloat16 else features.dtype)
        return {name: head(features, odt) for name, head in self.heads.items()}

    def forward(self, x):
        odt = self.model._odt(x)
        f = self.model._trunk(x)[0]
        return {"output": {name: head(f, odt) for name, head in self.heads.items()}}


# ----------------------------------------------------------------------------- builders
def _heads_cfg(cfg):
    raw = getattr(cfg.model, "heads", None)
    if not raw:
        return {}, None
    return dict(raw), getattr(cfg.model, "primary_head", None)


def _num_classes(cfg, head_cfg) -> int:
    # mednext_models.py:284-291
    if head_cfg:
        return max(1, sum(int(_cfg_value(s, "out_channels", 0)) for s in head_cfg.values()))
    return int(cfg.model.out_channels)


@register_architecture("mednext")
def build_mednext(cfg) -> ConnectomicsModel:
    """MedNeXt S/B/M/L (k in 3/5/7) on the B200 engine (``mednext_models.py:303-397``)."""
    mcfg = cfg.model.mednext
    size = getattr(mcfg, "size", "S")
    k = getattr(mcfg, "kernel_size", 3)
    ds = getattr(getattr(cfg.model, "loss", None), "deep_supervision", False)
    head_cfg, primary = _heads_cfg(cfg)
    if size not in ("S", "B", "M", "L"):
        raise ValueError(f"MedNeXt model_size must be 'S', 'B', 'M', or 'L'. Got: {size}\n"
                         "Model sizes:\n  - S (Small): 5.6M params\n  - B (Base): 10.5M params\n"
                         "  - M (Medium): 17.6M params\n  - L (Large): 61.8M params")
    if k not in (3, 5, 7):
        raise ValueError(f"MedNeXt kernel_size must be 3, 5, or 7. Got: {k}\nRecommended: Start with kernel_size=3")
    model = create_mednext_v1(cfg.model.in_channels, _num_classes(cfg, head_cfg), size, k, ds)
    style = getattr(mcfg, "checkpoint_style", None)
    if style is not None:
        if style != "outside_block":
            raise ValueError(f"model.mednext.checkpoint_style must be None or 'outside_block', got: {style!r}")
        model.outside_block_checkpointing = True
    if head_cfg:
        return MedNeXtMultiHeadWrapper(model, head_cfg, primary_head=primary)
    return MedNeXtWrapper(model, deep_supervision=ds)


@register_architecture("mednext_custom")
def build_mednext_custom(cfg) -> ConnectomicsModel:
    """MedNeXt with explicit architecture parameters (``mednext_models.py:400-483``)."""
    head_cfg, primary = _heads_cfg(cfg)
    m = cfg.model.mednext
    params = dict(
        in_channels=cfg.model.in_channels, n_channels=getattr(m, "base_channels", 32),
        n_classes=_num_classes(cfg, head_cfg), exp_r=getattr(m, "exp_r", 4),
        kernel_size=getattr(m, "kernel_size", 7),
        deep_supervision=getattr(cfg.model.loss, "deep_supervision", False),
        do_res=getattr(m, "do_res", True), do_res_up_down=getattr(m, "do_res_up_down", True),
        block_counts=getattr(m, "block_counts", [2] * 9), checkpoint_style=getattr(m, "checkpoint_style", None),
        norm_type=getattr(m, "norm", "group"), dim=getattr(m, "dim", "3d"), grn=getattr(m, "grn", False))
    if params["dim"] not in ("2d", "3d"):
        raise ValueError(f"mednext_dim must be '2d' or '3d', got: {params['dim']}")
    if params["norm_type"] not in ("group", "layer"):
        raise ValueError(f"mednext_norm must be 'group' or 'layer', got: {params['norm_type']}")
    if len(params["block_counts"]) != 9:
        raise ValueError("mednext_block_counts must have exactly 9 elements (one per level), "
                         f"got {len(params['block_counts'])}")
    if isinstance(params["exp_r"], (list, tuple)) or hasattr(params["exp_r"], "__iter__"):
        params["exp_r"] = [int(v) for v in params["exp_r"]]
    params["block_counts"] = [int(v) for v in params["block_counts"]]
    model = MedNeXt(**params)
    if head_cfg:
        return MedNeXtMultiHeadWrapper(model, head_cfg, primary_head=primary)
    return MedNeXtWrapper(model, deep_supervision=params["deep_supervision"])


def upkern_load_weights(target_model, source_model):
    """``mednext_models.py:487-537`` — initialise a large-kernel MedNeXt from a trained small-kernel one (UpKern).

    The reference delegates to ``nnunet_mednext.run.load_weights.upkern_load_weights`` (un-vendored); its published algorithm is
    restated here: every key present in both state dicts is copied when the spatial kernel dims agree and resized with
    ``F.interpolate(..., mode="trilinear")`` when they do not (channel dims must agree); keys missing from the source keep the
    target's initialisation.  Accepts the wrappers (``.model``) or bare ``MedNeXt`` modules and returns ``target_model``."""
    import torch.nn.functional as F
    tgt = getattr(target_model, "model", target_model)
    src = getattr(source_model, "model", source_model)
    src_sd, tgt_sd = src.state_dict(), tgt.state_dict()
    for key, cur in tgt_sd.items():
        if key not in src_sd:
            continue
        old = src_sd[key]
        if tuple(old.shape) == tuple(cur.shape):
            tgt_sd[key] = old.detach().clone().to(cur.device, cur.dtype)
            continue
        if old.dim() != cur.dim() or old.dim() < 3 or tuple(old.shape[:2]) != tuple(cur.shape[:2]):
            raise ValueError(f"UpKern: {key} has shape {tuple(old.shape)} in the source and {tuple(cur.shape)} in the target; "
                             "the models must have identical architecture except kernel size")
        tgt_sd[key] = F.interpolate(old.detach().float(), size=tuple(cur.shape[2:]), mode="trilinear").to(cur.device, cur.dtype)
    tgt.load_state_dict(tgt_sd)
    _lib_epoch_bump()
    return target_model


def _lib_epoch_bump() -> None:
    L.PARAM_EPOCH[0] += 1       # kernel-layout weight caches repack (parameter storage was rewritten wholesale)


__all__ = ["upkern_load_weights", "MedNeXt", "MedNeXtBlock", "MedNeXtDownBlock", "MedNeXtUpBlock", "OutBlock", "MedNeXtWrapper",
           "MedNeXtTaskHead", "MedNeXtMultiHeadWrapper", "build_mednext", "build_mednext_custom",
           "create_mednext_v1"]
