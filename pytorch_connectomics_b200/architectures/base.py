"""Model interface seam (drop-in for ``connectomics/models/architectures/base.py:17-87``)."""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Dict, Union

import torch
import torch.nn as nn


class ConnectomicsModel(nn.Module, ABC):
    """Forward contract: ``Tensor[B,ncls,D,H,W]`` or ``{"output", "ds_1".."ds_4"}`` (deep
    supervision) or ``{"output": {head: Tensor}}`` (multi-head)."""

    def __init__(self) -> None:
        super().__init__()
        self.supports_deep_supervision = False
        self.output_scales = 1

    @abstractmethod
    def forward(self, x: torch.Tensor) -> Union[torch.Tensor, Dict[str, torch.Tensor]]:
        raise NotImplementedError

    def get_model_info(self) -> Dict[str, Any]:
        params = list(self.parameters())
        return {
            "name": type(self).__name__,
            "deep_supervision": self.supports_deep_supervision,
            "output_scales": self.output_scales,
            "parameters": sum(p.numel() for p in params),
            "trainable_parameters": sum(p.numel() for p in params if p.requires_grad),
        }

    def __repr__(self) -> str:
        i = self.get_model_info()
        return f"{i['name']}(parameters={i['parameters']:,}, deep_supervision={i['deep_supervision']})"


__all__ = ["ConnectomicsModel"]
