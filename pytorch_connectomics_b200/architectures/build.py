"""``build_model(cfg)`` seam (drop-in for ``connectomics/models/build.py:24-72``)."""

from __future__ import annotations

import logging

from .registry import get_architecture_builder

logger = logging.getLogger(__name__)


def build_model(cfg):
    arch = cfg.model.arch.type
    model = get_architecture_builder(arch)(cfg)   # ValueError lists the available names
    logger.info("Model: %s (architecture: %s)", type(model).__name__, arch)
    if hasattr(model, "get_model_info"):
        info = model.get_model_info()
        logger.info("  Parameters: %s  Trainable: %s  Deep Supervision: %s",
                    f"{info['parameters']:,}", f"{info['trainable_parameters']:,}", info["deep_supervision"])
    return model


__all__ = ["build_model"]
