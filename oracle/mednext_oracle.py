"""CPU oracle: pure-torch restatement of the MedNeXt networks the reference builds.

TEST INFRASTRUCTURE — never imported by the product package.

The reference (``connectomics/models/architectures/mednext_models.py:23-32,374-380,479``)
does not contain the network arithmetic: it imports ``MedNeXt``, ``MedNeXtBlock`` and
``create_mednext_v1`` from the third-party package ``nnunet_mednext``
(``pip install git+https://github.com/PytorchConnectomics/MedNeXt.git`` — a fork of
MIC-DKFZ/MedNeXt, **no version pin**, ``INSTALLATION.md:157``).  That package is not
present under ``/root/reference`` and is not installed in this image, so this file
restates its published algorithm (upstream files
``nnunet_mednext/network_architecture/mednextv1/{blocks.py,MedNextV1.py,create_mednext_v1.py}``)
with the exact upstream attribute names / ``state_dict`` keys.

PARITY PINNING.  The restatement is pinned by what the reference itself constrains:
  * parameter counts quoted at ``mednext_models.py:309-312`` (5.6M/10.5M/17.6M/61.8M
    for k=3; 5.9/11.0/18.3/63.0 for k=5) — ``tests/test_oracle_mednext.py``;
  * the attributes the reference introspects (``mednext_models.py:99-126,215-231``):
    ``dec_block_0[0].conv1.kernel_size``, ``.conv2.{in,out}_channels``, ``.norm``,
    ``.do_res``, ``.dim``, ``.grn``, ``do_ds``, ``forward_features``, ``stem.out_channels``,
    ``outside_block_checkpointing``;
  * the identities in the reference's own tests
    (``tests/unit/test_mednext_features.py:26-55``): ``forward_output(forward_features(x))
    == model(x)``, deep supervision returns a 5-list.
  * one independent numeric check: the block skeleton with the channels-first LayerNorm equals torchvision's ConvNeXt block
    (``torchvision.models.convnext.CNBlock``, layer scale 1, eps 1e-5) from the same weights —
    ``tests/test_oracle_mednext.py::test_block_restatement_equals_torchvisions_convnext_block``.
Numeric outputs of ``nnunet_mednext`` itself are NOT available here: for the network
arithmetic this oracle is "parity unpinned" against the third-party wheel (stated in
DESIGN.md); every op is a stock ``torch.nn.functional`` call, which is what upstream runs.
"""

from __future__ import annotations

from typing import List, Sequence, Union

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.checkpoint import checkpoint

# ----------------------------------------------------------------------------- rounding-matched mode
# The B200 engine stores every inter-kernel activation (and every inter-kernel gradient) as bf16 and feeds the
# tensor cores bf16 pointwise weights with fp32 accumulation.  `with bf16_matched():` makes this oracle round at
# exactly those points (same stock torch ops otherwise, fp32 arithmetic between the rounding points), so that the
# engine can be held to a bound far below the bf16 storage error itself (tests: rel-L2 <= 2e-3 and a max-abs bound)
# instead of "not worse than 1.5x the autocast path".  Rounding points:
#   forward : stem output, conv1 output y, normalised y (GEMM A operand), GELU output (second GEMM's A operand),
#             block output after bias + residual / res_conv / skip; conv2/conv3/res_conv weights;
#   backward: the gradients of the same tensors (dOut, dh = dG*GELU', dYhat, dy, dx) — dG stays fp32 (TMEM).
_MATCH = False


class bf16_matched:
    def __enter__(self):
        global _MATCH
        self._prev, _MATCH = _MATCH, True
        return self

    def __exit__(self, *exc):
        global _MATCH
        _MATCH = self._prev
        return False


def _r(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


class _Q(torch.autograd.Function):
    """round to bf16 in forward (fwd=True) and round the gradient to bf16 in backward (bwd=True)."""

    @staticmethod
    def forward(ctx, x, fwd, bwd):
        ctx.bwd = bwd
        return _r(x) if fwd else x.clone()

    @staticmethod
    def backward(ctx, g):
        return (_r(g) if ctx.bwd else g), None, None


def q(x, fwd=True, bwd=True):
    return _Q.apply(x, fwd, bwd)


def qw(w):
    """bf16-rounded weight with a straight-through gradient (the master weight stays fp32)."""
    return w + (_r(w) - w).detach()


class MedNeXtBlock(nn.Module):
    """upstream blocks.py::MedNeXtBlock (3-D / 2-D, GroupNorm|LayerNorm, optional GRN)."""

    def __init__(self, in_channels: int, out_channels: int, exp_r: int = 4, kernel_size: int = 7,
                 do_res: bool = True, norm_type: str = "group", n_groups=None, dim: str = "3d",
                 grn: bool = False):
        super().__init__()
        self.do_res = do_res
        assert dim in ("2d", "3d")
        self.dim = dim
        conv = nn.Conv2d if dim == "2d" else nn.Conv3d
        self.conv1 = conv(in_channels, in_channels, kernel_size=kernel_size, stride=1,
                          padding=kernel_size // 2,
                          groups=in_channels if n_groups is None else n_groups)
        if norm_type == "group":
            self.norm = nn.GroupNorm(num_groups=in_channels, num_channels=in_channels)
        elif norm_type == "layer":
            self.norm = _ChannelsFirstLayerNorm(in_channels)
        else:
            raise ValueError(norm_type)
        self.conv2 = conv(in_channels, exp_r * in_channels, kernel_size=1, stride=1, padding=0)
        self.act = nn.GELU()
        self.conv3 = conv(exp_r * in_channels, out_channels, kernel_size=1, stride=1, padding=0)
        self.grn = grn
        if grn:
            shape = (1, exp_r * in_channels) + (1,) * (3 if dim == "3d" else 2)
            self.grn_beta = nn.Parameter(torch.zeros(shape), requires_grad=True)
            self.grn_gamma = nn.Parameter(torch.zeros(shape), requires_grad=True)

    def _conv_q(self, conv, x):
        """the module's conv / transposed conv with bf16-rounded weights (bias stays fp32)"""
        fn = {nn.Conv3d: F.conv3d, nn.Conv2d: F.conv2d, nn.ConvTranspose3d: F.conv_transpose3d,
              nn.ConvTranspose2d: F.conv_transpose2d}[type(conv)]
        return fn(x, qw(conv.weight), conv.bias, stride=conv.stride, padding=conv.padding)

    def _core_matched(self, x):
        """block body with the engine's rounding points; returns the UNROUNDED conv3 output (the caller adds the
        residual / res_conv / skip in fp32 and rounds once, as the fused epilogue does)."""
        y = q(self.conv1(x))                                  # depthwise: fp32 taps, bf16 in/out
        a = q(self.norm(y))                                   # GEMM A operand
        h = q(self._conv_q(self.conv2, a), fwd=False)         # fp32 accumulator; its gradient dh is stored bf16
        g = q(self.act(h), bwd=False)                         # GELU output -> bf16 smem tile; dG stays fp32 (TMEM)
        if self.grn:
            dims = (-3, -2, -1) if self.dim == "3d" else (-2, -1)
            gx = torch.norm(g, p=2, dim=dims, keepdim=True)
            nx = gx / (gx.mean(dim=1, keepdim=True) + 1e-6)
            g = q(self.grn_gamma * (g * nx) + self.grn_beta + g, bwd=False)
        return self._conv_q(self.conv3, g)

    def _core(self, x):
        y = self.conv1(x)
        y = self.act(self.conv2(self.norm(y)))
        if self.grn:
            dims = (-3, -2, -1) if self.dim == "3d" else (-2, -1)
            gx = torch.norm(y, p=2, dim=dims, keepdim=True)
            nx = gx / (gx.mean(dim=1, keepdim=True) + 1e-6)
            y = self.grn_gamma * (y * nx) + self.grn_beta + y
        return self.conv3(y)

    def forward(self, x, dummy_tensor=None):
        if _MATCH and type(self) is MedNeXtBlock:
            y = self._core_matched(x)
            return q(x + y if self.do_res else y)
        y = self._core(x)
        return x + y if self.do_res else y


class _ChannelsFirstLayerNorm(nn.Module):
    """upstream blocks.py::LayerNorm(data_format='channels_first'), eps 1e-5."""

    def __init__(self, c, eps=1e-5):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.eps = eps

    def forward(self, x):
        u = x.mean(1, keepdim=True)
        s = (x - u).pow(2).mean(1, keepdim=True)
        x = (x - u) / torch.sqrt(s + self.eps)
        shape = (1, -1) + (1,) * (x.dim() - 2)
        return self.weight.view(shape) * x + self.bias.view(shape)


class MedNeXtDownBlock(MedNeXtBlock):
    """upstream blocks.py::MedNeXtDownBlock — stride-2 depthwise conv1, strided 1x1 residual."""

    def __init__(self, in_channels, out_channels, exp_r=4, kernel_size=7, do_res=False,
                 norm_type="group", dim="3d", grn=False):
        super().__init__(in_channels, out_channels, exp_r, kernel_size, do_res=False,
                         norm_type=norm_type, dim=dim, grn=grn)
        conv = nn.Conv2d if dim == "2d" else nn.Conv3d
        self.resample_do_res = do_res
        if do_res:
            self.res_conv = conv(in_channels, out_channels, kernel_size=1, stride=2)
        self.conv1 = conv(in_channels, in_channels, kernel_size=kernel_size, stride=2,
                          padding=kernel_size // 2, groups=in_channels)

    def forward(self, x, dummy_tensor=None):
        if _MATCH:
            y = self._core_matched(x)
            if self.resample_do_res:
                y = y + self._conv_q(self.res_conv, x)
            return q(y)
        y = super().forward(x)
        if self.resample_do_res:
            y = y + self.res_conv(x)
        return y


class MedNeXtUpBlock(MedNeXtBlock):
    """upstream blocks.py::MedNeXtUpBlock — transposed stride-2 depthwise conv1, the block output
    (spatial 2s-1) is zero-padded by one voxel at the FRONT of every spatial axis."""

    def __init__(self, in_channels, out_channels, exp_r=4, kernel_size=7, do_res=False,
                 norm_type="group", dim="3d", grn=False):
        super().__init__(in_channels, out_channels, exp_r, kernel_size, do_res=False,
                         norm_type=norm_type, dim=dim, grn=grn)
        self.resample_do_res = do_res
        convt = nn.ConvTranspose2d if dim == "2d" else nn.ConvTranspose3d
        if do_res:
            self.res_conv = convt(in_channels, out_channels, kernel_size=1, stride=2)
        self.conv1 = convt(in_channels, in_channels, kernel_size=kernel_size, stride=2,
                           padding=kernel_size // 2, groups=in_channels)

    def forward(self, x, dummy_tensor=None, skip=None):
        pad = (1, 0, 1, 0, 1, 0) if self.dim == "3d" else (1, 0, 1, 0)
        if _MATCH:      # the engine adds res_conv and the encoder skip in the same fp32 epilogue, then rounds once
            y = F.pad(self._core_matched(x), pad)
            if self.resample_do_res:
                y = y + F.pad(self._conv_q(self.res_conv, x), pad)
            return q(y if skip is None else skip + y)
        y = super().forward(x)
        y = F.pad(y, pad)
        if self.resample_do_res:
            y = y + F.pad(self.res_conv(x), pad)
        return y if skip is None else skip + y


class OutBlock(nn.Module):
    """upstream blocks.py::OutBlock — ConvTranspose(k=1)."""

    def __init__(self, in_channels, n_classes, dim="3d"):
        super().__init__()
        convt = nn.ConvTranspose2d if dim == "2d" else nn.ConvTranspose3d
        self.conv_out = convt(in_channels, n_classes, kernel_size=1)

    def forward(self, x, dummy_tensor=None):
        return self.conv_out(x)


class MedNeXt(nn.Module):
    """upstream MedNextV1.py::MedNeXt + the fork's forward_features/forward_output
    (pinned by reference tests/unit/test_mednext_features.py:26-39)."""

    def __init__(self, in_channels: int, n_channels: int, n_classes: int,
                 exp_r: Union[int, Sequence[int]] = 4, kernel_size: int = 7,
                 enc_kernel_size: int = None, dec_kernel_size: int = None,
                 deep_supervision: bool = False, do_res: bool = False,
                 do_res_up_down: bool = False, checkpoint_style: str = None,
                 block_counts: Sequence[int] = (2, 2, 2, 2, 2, 2, 2, 2, 2),
                 norm_type: str = "group", dim: str = "3d", grn: bool = False):
        super().__init__()
        self.do_ds = deep_supervision
        assert checkpoint_style in (None, "outside_block")
        self.inside_block_checkpointing = False
        self.outside_block_checkpointing = checkpoint_style == "outside_block"
        assert dim in ("2d", "3d")
        if kernel_size is not None:
            enc_kernel_size = kernel_size
            dec_kernel_size = kernel_size
        conv = nn.Conv2d if dim == "2d" else nn.Conv3d
        self.stem = conv(in_channels, n_channels, kernel_size=1)
        if isinstance(exp_r, int):
            exp_r = [exp_r] * len(block_counts)
        exp_r = list(exp_r)
        n = n_channels
        kw = dict(norm_type=norm_type, dim=dim, grn=grn)

        def stage(c, i, k):
            return nn.Sequential(*[
                MedNeXtBlock(c, c, exp_r[i], k, do_res=do_res, **kw)
                for _ in range(block_counts[i])])

        self.enc_block_0 = stage(n, 0, enc_kernel_size)
        self.down_0 = MedNeXtDownBlock(n, 2 * n, exp_r[1], enc_kernel_size, do_res=do_res_up_down, **kw)
        self.enc_block_1 = stage(2 * n, 1, enc_kernel_size)
        self.down_1 = MedNeXtDownBlock(2 * n, 4 * n, exp_r[2], enc_kernel_size, do_res=do_res_up_down, **kw)
        self.enc_block_2 = stage(4 * n, 2, enc_kernel_size)
        self.down_2 = MedNeXtDownBlock(4 * n, 8 * n, exp_r[3], enc_kernel_size, do_res=do_res_up_down, **kw)
        self.enc_block_3 = stage(8 * n, 3, enc_kernel_size)
        self.down_3 = MedNeXtDownBlock(8 * n, 16 * n, exp_r[4], enc_kernel_size, do_res=do_res_up_down, **kw)
        self.bottleneck = stage(16 * n, 4, dec_kernel_size)
        self.up_3 = MedNeXtUpBlock(16 * n, 8 * n, exp_r[5], dec_kernel_size, do_res=do_res_up_down, **kw)
        self.dec_block_3 = stage(8 * n, 5, dec_kernel_size)
        self.up_2 = MedNeXtUpBlock(8 * n, 4 * n, exp_r[6], dec_kernel_size, do_res=do_res_up_down, **kw)
        self.dec_block_2 = stage(4 * n, 6, dec_kernel_size)
        self.up_1 = MedNeXtUpBlock(4 * n, 2 * n, exp_r[7], dec_kernel_size, do_res=do_res_up_down, **kw)
        self.dec_block_1 = stage(2 * n, 7, dec_kernel_size)
        self.up_0 = MedNeXtUpBlock(2 * n, n, exp_r[8], dec_kernel_size, do_res=do_res_up_down, **kw)
        self.dec_block_0 = stage(n, 8, dec_kernel_size)
        self.out_0 = OutBlock(n, n_classes, dim=dim)
        # dummy tensor keeps torch.utils.checkpoint happy when no input requires grad
        self.dummy_tensor = nn.Parameter(torch.tensor([1.0]), requires_grad=True)
        if deep_supervision:
            self.out_1 = OutBlock(2 * n, n_classes, dim=dim)
            self.out_2 = OutBlock(4 * n, n_classes, dim=dim)
            self.out_3 = OutBlock(8 * n, n_classes, dim=dim)
            self.out_4 = OutBlock(16 * n, n_classes, dim=dim)
        self.block_counts = list(block_counts)

    def _run(self, mod, x):
        if self.outside_block_checkpointing and torch.is_grad_enabled():
            if isinstance(mod, nn.Sequential):
                for layer in mod:
                    x = checkpoint(layer, x, self.dummy_tensor, use_reentrant=False)
                return x
            return checkpoint(mod, x, self.dummy_tensor, use_reentrant=False)
        return mod(x)

    def _trunk(self, x) -> List[torch.Tensor]:
        """Returns [features_level0, dec1, dec2, dec3, bottleneck] (inputs of out_0..out_4)."""
        x = self.stem(x)
        if _MATCH:
            return self._trunk_matched(q(x))
        r0 = self._run(self.enc_block_0, x)
        x = self._run(self.down_0, r0)
        r1 = self._run(self.enc_block_1, x)
        x = self._run(self.down_1, r1)
        r2 = self._run(self.enc_block_2, x)
        x = self._run(self.down_2, r2)
        r3 = self._run(self.enc_block_3, x)
        x = self._run(self.down_3, r3)
        b = self._run(self.bottleneck, x)
        x = self._run(self.dec_block_3, r3 + self._run(self.up_3, b))
        d3 = x
        x = self._run(self.dec_block_2, r2 + self._run(self.up_2, x))
        d2 = x
        x = self._run(self.dec_block_1, r1 + self._run(self.up_1, x))
        d1 = x
        x = self._run(self.dec_block_0, r0 + self._run(self.up_0, x))
        return [x, d1, d2, d3, b]

    def _trunk_matched(self, x) -> List[torch.Tensor]:
        """same dataflow with the encoder skip handed to the up block (one rounding after the fused add)"""
        r0 = self.enc_block_0(x)
        r1 = self.enc_block_1(self.down_0(r0))
        r2 = self.enc_block_2(self.down_1(r1))
        r3 = self.enc_block_3(self.down_2(r2))
        b = self.bottleneck(self.down_3(r3))
        d3 = self.dec_block_3(self.up_3(b, skip=r3))
        d2 = self.dec_block_2(self.up_2(d3, skip=r2))
        d1 = self.dec_block_1(self.up_1(d2, skip=r1))
        f0 = self.dec_block_0(self.up_0(d1, skip=r0))
        return [f0, d1, d2, d3, b]

    def forward_features(self, x):
        return self._trunk(x)[0]

    def forward_output(self, features):
        return self.out_0(features)

    def forward(self, x):
        f = self._trunk(x)
        y = self.out_0(f[0])
        if self.do_ds:
            return [y, self.out_1(f[1]), self.out_2(f[2]), self.out_3(f[3]), self.out_4(f[4])]
        return y


MEDNEXT_V1_TABLE = {
    # upstream create_mednext_v1.py (n_channels=32, do_res=True, do_res_up_down=True)
    "S": dict(exp_r=2, block_counts=[2] * 9, checkpoint_style=None),
    "B": dict(exp_r=[2, 3, 4, 4, 4, 4, 4, 3, 2], block_counts=[2] * 9, checkpoint_style=None),
    "M": dict(exp_r=[2, 3, 4, 4, 4, 4, 4, 3, 2], block_counts=[3, 4, 4, 4, 4, 4, 4, 4, 3],
              checkpoint_style="outside_block"),
    "L": dict(exp_r=[3, 4, 8, 8, 8, 8, 8, 4, 3], block_counts=[3, 4, 8, 8, 8, 8, 8, 4, 3],
              checkpoint_style="outside_block"),
}


def create_mednext_v1(num_input_channels, num_classes, model_id, kernel_size=3,
                      deep_supervision=False):
    spec = MEDNEXT_V1_TABLE[model_id]
    return MedNeXt(in_channels=num_input_channels, n_channels=32, n_classes=num_classes,
                   exp_r=spec["exp_r"], kernel_size=kernel_size,
                   deep_supervision=deep_supervision, do_res=True, do_res_up_down=True,
                   block_counts=spec["block_counts"], checkpoint_style=spec["checkpoint_style"])
