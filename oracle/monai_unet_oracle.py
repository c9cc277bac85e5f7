"""CPU oracle: pure-torch restatement of the MONAI ``UNet`` the reference builds for ``monai_unet``.

TEST INFRASTRUCTURE — never imported by the product package.

The reference (``connectomics/models/architectures/monai_models.py:197-250``) instantiates
``monai.networks.nets.UNet`` (through ``UpsampleModeUNet``, ``:84-139``; default ``upsample_mode="deconv"``
leaves MONAI's own up layer) with ``spatial_dims, in_channels, out_channels, channels=filters,
strides=[2]*(len-1), num_res_units, kernel_size, norm, dropout``.  ``monai>=0.9.1`` is un-pinned
(``pyproject.toml:51``), not under ``/root/reference`` and not installed here, so this file restates the
published modules (``monai/networks/nets/unet.py``, ``blocks/convolutions.py`` — ``Convolution``,
``ResidualUnit`` —, ``blocks/acti_norm.py`` — ``ADN`` —, ``layers/simplelayers.py`` — ``SkipConnection``)
with the upstream child-module names, so ``state_dict`` keys match (``model.0.conv.unit0.conv.weight``,
``...adn.N.*``, ``...adn.A.weight``, ``model.0.residual.*``, ``model.1.submodule.*``, ``model.2.0.conv.*`` …).
Arguments the reference does not pass keep MONAI's defaults: ``act=PReLU`` (one slope, 0.25), ``up_kernel_size=3``,
``bias=True``, ``adn_ordering="NDA"``.

PARITY PINNING: numeric outputs of the MONAI wheel are not available offline — "parity unpinned" against
MONAI itself; pinned by structure only (key names / shapes from SURVEY Appendix A.2, the shape contract of
``tests/integration/test_e2e_training.py:132-187``).  Every op is the stock torch op MONAI calls.
"""

from __future__ import annotations

from typing import Sequence, Tuple, Union

import torch
import torch.nn as nn


def _norm_layer(norm, channels: int, dims: int = 3) -> nn.Module:
    if isinstance(norm, (tuple, list)):
        name, kw = norm[0], dict(norm[1])
    else:
        name, kw = norm, {}
    name = str(name).lower()
    if name == "batch":
        return (nn.BatchNorm3d if dims == 3 else nn.BatchNorm2d)(channels, **kw)
    if name == "group":
        return nn.GroupNorm(num_channels=channels, **kw)
    if name == "instance":
        return (nn.InstanceNorm3d if dims == 3 else nn.InstanceNorm2d)(channels, **kw)
    raise ValueError(f"unsupported norm {norm!r}")


class ADN(nn.Sequential):
    """monai.networks.blocks.ADN with ordering "NDA": norm -> dropout -> activation (children N, D, A)."""

    def __init__(self, channels: int, norm, dropout, dims: int = 3):
        super().__init__()
        self.add_module("N", _norm_layer(norm, channels, dims))
        if dropout is not None:
            self.add_module("D", (nn.Dropout3d if False else nn.Dropout)(float(dropout)))
        self.add_module("A", nn.PReLU())


class Convolution(nn.Sequential):
    """monai.networks.blocks.Convolution: (Conv | ConvTranspose) [+ ADN]; same padding; output_padding = stride-1."""

    def __init__(self, dims, in_channels, out_channels, strides=1, kernel_size=3, norm="batch", dropout=0.0, bias=True,
                 conv_only=False, is_transposed=False):
        super().__init__()
        pad = (kernel_size - 1) // 2
        if dims != 3:
            raise ValueError("oracle restates the 3-D variant only")
        if is_transposed:
            conv = nn.ConvTranspose3d(in_channels, out_channels, kernel_size, stride=strides, padding=pad,
                                      output_padding=strides - 1, bias=bias)
        else:
            conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride=strides, padding=pad, bias=bias)
        self.add_module("conv", conv)
        if not conv_only:
            self.add_module("adn", ADN(out_channels, norm, dropout, dims))


class ResidualUnit(nn.Module):
    """monai.networks.blocks.ResidualUnit: conv = unit0..unit{n-1}; residual = strided k-conv | 1x1 conv | identity."""

    def __init__(self, dims, in_channels, out_channels, strides=1, kernel_size=3, subunits=2, norm="batch", dropout=0.0,
                 bias=True, last_conv_only=False):
        super().__init__()
        self.conv = nn.Sequential()
        self.residual: nn.Module = nn.Identity()
        subunits = max(1, subunits)
        sch, sst = in_channels, strides
        for su in range(subunits):
            only = last_conv_only and su == subunits - 1
            self.conv.add_module(f"unit{su:d}", Convolution(dims, sch, out_channels, sst, kernel_size, norm, dropout, bias, only))
            sch, sst = out_channels, 1
        if strides != 1 or in_channels != out_channels:
            rk, rp = (kernel_size, (kernel_size - 1) // 2) if strides != 1 else (1, 0)
            self.residual = nn.Conv3d(in_channels, out_channels, rk, strides, rp, bias=bias)

    def forward(self, x):
        return self.conv(x) + self.residual(x)


class SkipConnection(nn.Module):
    def __init__(self, submodule: nn.Module):
        super().__init__()
        self.submodule = submodule

    def forward(self, x):
        return torch.cat([x, self.submodule(x)], dim=1)


class UNet(nn.Module):
    """monai.networks.nets.UNet (3-D), as parameterised by monai_models.py:235-248."""

    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, channels: Sequence[int], strides: Sequence[int],
                 kernel_size: int = 3, up_kernel_size: int = 3, num_res_units: int = 0,
                 norm: Union[str, Tuple] = "batch", dropout: float = 0.0, bias: bool = True):
        super().__init__()
        if len(channels) < 2:
            raise ValueError("the length of `channels` should be no less than 2.")
        if len(strides) < len(channels) - 1:
            raise ValueError("the length of `strides` should equal to `len(channels) - 1`.")
        self.dimensions, self.kernel_size, self.up_kernel_size = spatial_dims, kernel_size, up_kernel_size
        self.num_res_units, self.norm, self.dropout, self.bias = num_res_units, norm, dropout, bias

        def block(inc, outc, ch, st, is_top):
            c, s = ch[0], st[0]
            if len(ch) > 2:
                sub, upc = block(c, c, ch[1:], st[1:], False), c * 2
            else:
                sub, upc = self._down(c, ch[1], 1), c + ch[1]
            return nn.Sequential(self._down(inc, c, s), SkipConnection(sub), self._up(upc, outc, s, is_top))

        self.model = block(in_channels, out_channels, list(channels), list(strides), True)

    def _down(self, i, o, s):
        if self.num_res_units > 0:
            return ResidualUnit(self.dimensions, i, o, s, self.kernel_size, self.num_res_units, self.norm, self.dropout, self.bias)
        return Convolution(self.dimensions, i, o, s, self.kernel_size, self.norm, self.dropout, self.bias)

    def _up(self, i, o, s, is_top):
        conv = Convolution(self.dimensions, i, o, s, self.up_kernel_size, self.norm, self.dropout, self.bias,
                           conv_only=is_top and self.num_res_units == 0, is_transposed=True)
        if self.num_res_units > 0:
            ru = ResidualUnit(self.dimensions, o, o, 1, self.kernel_size, 1, self.norm, self.dropout, self.bias,
                              last_conv_only=is_top)
            return nn.Sequential(conv, ru)
        return conv

    def forward(self, x):
        return self.model(x)
