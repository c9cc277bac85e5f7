"""autograd.Function wrappers for the MedNeXt ops (backward kernels: csrc/mednext_bwd.cu)."""

from __future__ import annotations

import torch


class _NotYet(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *a):
        raise NotImplementedError("pcb200: MedNeXt backward kernels are not built yet; run under torch.no_grad()")


StemFn = BlockFn = HeadFn = _NotYet
