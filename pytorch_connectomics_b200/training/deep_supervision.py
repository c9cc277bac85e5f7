"""Deep-supervision side of the training step: what ``connectomics/training/losses/orchestrator.py:817-950``
does AROUND the per-scale loss when a MedNeXt trunk returns ``{"output", "ds_1" .. "ds_4"}`` (or the 5-list) — the scale weights
and the resize of the target to each head's resolution.  The per-scale loss itself stays the caller's (losses are outside the
hot path); this module only makes ``ArenaTrainStep(model, partial(deep_supervision_loss, loss_fn=...), ...)`` reproduce the
reference's weighted sum.  Targets are resized with device tensor ops (one interpolate per scale on a 1-to-few-channel label
volume: < 1 % of the step's bytes)."""

from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn.functional as F

_LABEL_DTYPES = (torch.long, torch.int, torch.int32, torch.int64, torch.uint8)


def deep_supervision_weights(n_outputs: int, configured: Optional[Sequence[float]] = None) -> List[float]:
    """``orchestrator.py:834-847``: ``loss.deep_supervision_weights`` when it names every output, else 1, 1/2, 1/4, ...
    (a list that is too SHORT is dropped as a whole — the reference warns and falls back, it does not pad)."""
    if configured is not None and len(configured) >= n_outputs:
        return [float(w) for w in configured]
    return [0.5 ** i for i in range(n_outputs)]


def match_target_to_output(target: torch.Tensor, output: torch.Tensor) -> torch.Tensor:
    """``orchestrator.py:879-950``: the target at a head's resolution.  Integer labels: nearest neighbour (through fp32, back to
    int64).  Continuous targets: trilinear, ``align_corners=False``, then clamped to [-1, 1] when the ORIGINAL target lies in
    [-1.5, 1.5] (interpolation overshoot of tanh / sigmoid-range targets); the reference's second band ([0, 1.5] -> [0, 1]) sits
    behind the first and can never fire, so it is not restated."""
    if target.shape == output.shape:
        return target
    size = tuple(int(v) for v in output.shape[2:])
    if target.dtype in _LABEL_DTYPES:
        return F.interpolate(target.float(), size=size, mode="nearest").long()
    resized = F.interpolate(target, size=size, mode="trilinear", align_corners=False)
    if bool(target.min() >= -1.5) and bool(target.max() <= 1.5):
        resized = resized.clamp(-1.0, 1.0)
    return resized


def split_outputs(outputs: Union[torch.Tensor, Sequence[torch.Tensor], Dict[str, Any]]) -> List[torch.Tensor]:
    """main output first, then ``ds_1`` .. ``ds_4`` in order (``orchestrator.py:831-832``); a bare tensor or the trunk's list pass"""
    if isinstance(outputs, dict):
        return [outputs["output"]] + [outputs[f"ds_{i}"] for i in range(1, 5) if f"ds_{i}" in outputs]
    if isinstance(outputs, (list, tuple)):
        return list(outputs)
    return [outputs]


def deep_supervision_loss(outputs, labels: torch.Tensor, loss_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], *,
                          weights: Optional[Sequence[float]] = None, return_terms: bool = False):
    """``sum_i w_i * loss_fn(output_i, match_target_to_output(labels, output_i))`` (``orchestrator.py:849-867``)"""
    scales = split_outputs(outputs)
    w = deep_supervision_weights(len(scales), weights)
    terms: List[Tuple[float, torch.Tensor]] = []
    total = None
    for out, wi in zip(scales, w):
        term = loss_fn(out, match_target_to_output(labels, out))
        terms.append((wi, term))
        total = term * wi if total is None else total + term * wi
    return (total, terms) if return_terms else total


__all__ = ["deep_supervision_loss", "deep_supervision_weights", "match_target_to_output", "split_outputs"]
