"""The two array transforms chunked inference applies to a chunk before it is written (``connectomics/inference/output.py:
145-243``): ``inference.prediction_transform`` (intensity scale + dtype) and ``inference.save_dtype``.  Host numpy on the
chunk that already left the GPU — the file writers, filename resolution and decoding around them are outside this path."""

from __future__ import annotations

import logging
from typing import Any, Optional

import numpy as np

logger = logging.getLogger(__name__)

_DTYPES = {name: getattr(np, name) for name in ("uint8", "int8", "uint16", "int16", "uint32", "int32", "float16", "float32",
                                                "float64")}


def _convert_intensity_dtype(data: np.ndarray, target: Optional[str], *, config_name: str) -> np.ndarray:
    """``output.py:145-185`` — integer targets clip to the type's range first; unknown names keep the dtype (warning)."""
    if target is None:
        return data
    if target not in _DTYPES:
        logger.warning("Unknown dtype '%s' in %s. Supported: %s. Keeping current dtype.", target, config_name, list(_DTYPES))
        return data
    dt = _DTYPES[target]
    if np.issubdtype(dt, np.integer):
        info = np.iinfo(dt)
        data = np.clip(data, info.min, info.max)
    return data.astype(dt, copy=False)


def _apply_intensity_transform(data: np.ndarray, *, intensity_scale, intensity_dtype, config_name: str) -> np.ndarray:
    """``output.py:188-210`` — a negative (or missing) scale disables scaling; scaling happens in float32."""
    if intensity_scale is not None and intensity_scale >= 0:
        data = data.astype(np.float32, copy=False)
        if intensity_scale != 1.0:
            data = data * float(intensity_scale)
    return _convert_intensity_dtype(data, intensity_dtype, config_name=config_name)


def apply_prediction_transform(cfg: Any, data: np.ndarray) -> np.ndarray:
    """``output.py:213-227``"""
    tc = getattr(getattr(cfg, "inference", None), "prediction_transform", None)
    if tc is None or not getattr(tc, "enabled", False):
        return data
    return _apply_intensity_transform(data, intensity_scale=getattr(tc, "intensity_scale", -1.0),
                                      intensity_dtype=getattr(tc, "intensity_dtype", None),
                                      config_name="inference.prediction_transform")


def apply_storage_dtype_transform(cfg: Any, data: np.ndarray) -> np.ndarray:
    """``output.py:230-243``"""
    target = getattr(getattr(cfg, "inference", None), "save_dtype", None)
    if target is None:
        return data
    return _convert_intensity_dtype(data, target, config_name="inference.save_dtype")


__all__ = ["apply_prediction_transform", "apply_storage_dtype_transform"]
