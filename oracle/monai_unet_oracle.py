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
                 bias=True, last_conv_only=False, act="PRELU", adn_ordering="NDA"):
        super().__init__()
        if str(act).upper() != "PRELU" or adn_ordering != "NDA":
            raise ValueError("oracle restates MONAI's defaults only: act='PRELU', adn_ordering='NDA'")
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


class UpSample(nn.Sequential):
    """monai.networks.blocks.UpSample in its ``mode="nontrainable"`` form (the one ``UpsampleModeUNet`` is for,
    ``monai_models.py:84-139``): ``preconv`` = Conv(in, out, kernel 1) when the channel counts differ, then
    ``upsample_non_trainable`` = ``nn.Upsample(scale_factor, interp mode, align_corners)``; the linear family resolves to the
    mode of the dimensionality (3-D: trilinear).  No norm / activation.  Restated from MONAI's published block — like the rest
    of this file, un-pinned against the wheel."""

    def __init__(self, spatial_dims, in_channels=None, out_channels=None, scale_factor=2, mode="deconv", interp_mode="linear",
                 align_corners=True, bias=True, **_unused):
        super().__init__()
        if spatial_dims != 3:
            raise ValueError("oracle restates the 3-D variant only")
        if str(mode).lower() != "nontrainable":
            raise NotImplementedError(f"oracle UpSample restates mode='nontrainable' only, got {mode!r}")
        out_channels = in_channels if out_channels is None else out_channels
        if out_channels != in_channels:
            self.add_module("preconv", nn.Conv3d(in_channels, out_channels, kernel_size=1, bias=bias))
        interp = str(getattr(interp_mode, "value", interp_mode)).lower()
        if interp in ("linear", "bilinear", "trilinear"):
            interp = "trilinear"
        sf = tuple(scale_factor) if isinstance(scale_factor, (tuple, list)) else (scale_factor,) * 3
        self.add_module("upsample_non_trainable", nn.Upsample(scale_factor=tuple(float(v) for v in sf), mode=interp,
                                                              align_corners=align_corners))


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
        self.act, self.adn_ordering = "PRELU", "NDA"            # MONAI's defaults, read by subclasses that build their own layers

        def block(inc, outc, ch, st, is_top):
            c, s = ch[0], st[0]
            if len(ch) > 2:
                sub, upc = block(c, c, ch[1:], st[1:], False), c * 2
            else:
                sub, upc = self._get_bottom_layer(c, ch[1]), c + ch[1]
            return nn.Sequential(self._get_down_layer(inc, c, s, is_top), SkipConnection(sub), self._get_up_layer(upc, outc, s, is_top))

        self.model = block(in_channels, out_channels, list(channels), list(strides), True)

    # MONAI's method names: a subclass (the reference's UpsampleModeUNet) overrides _get_up_layer
    def _get_down_layer(self, in_channels, out_channels, strides, is_top):
        return self._down(in_channels, out_channels, strides)

    def _get_bottom_layer(self, in_channels, out_channels):
        return self._down(in_channels, out_channels, 1)

    def _get_up_layer(self, in_channels, out_channels, strides, is_top):
        return self._up(in_channels, out_channels, strides, is_top)

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


class UpsampleModeUNet(UNet):
    """``connectomics/models/architectures/monai_models.py:84-139`` restated: MONAI's UNet whose up layers are
    ``UpSample(mode=upsample_mode)`` [+ a one-subunit ResidualUnit when ``num_res_units > 0``] unless the mode is "deconv".
    (Travels to the GPU box, where the reference itself is absent; ``tests/test_builders_vs_reference.py`` holds it against the
    REAL class executed in place.)"""

    def __init__(self, upsample_mode: str = "deconv", upsample_interp_mode: str = "linear", upsample_align_corners: bool = True,
                 **kwargs):
        self.upsample_mode, self.upsample_interp_mode = upsample_mode, upsample_interp_mode
        self.upsample_align_corners = upsample_align_corners
        super().__init__(**kwargs)

    def _get_up_layer(self, in_channels, out_channels, strides, is_top):
        if not self.upsample_mode or self.upsample_mode == "deconv":
            return super()._get_up_layer(in_channels, out_channels, strides, is_top)
        up = UpSample(self.dimensions, in_channels, out_channels, strides, mode=self.upsample_mode,
                      interp_mode=self.upsample_interp_mode, align_corners=self.upsample_align_corners, bias=self.bias)
        if self.num_res_units == 0:
            return up
        return nn.Sequential(up, ResidualUnit(self.dimensions, out_channels, out_channels, 1, self.kernel_size, 1, self.norm,
                                              self.dropout, self.bias, last_conv_only=is_top))
