"""CPU oracle of the test-time-augmentation path (TEST INFRASTRUCTURE ONLY — imported by tests/, smoke() and the
golden generator, never by the product).

Restates, for fully valid channels, ``connectomics/inference/tta.py:691-771`` (``_run_ensemble``: flip + rot90 view,
network, ``tta_affinity.py:364-369`` ``invert_view``), ``tta.py:312-402`` (``apply_preprocessing``) and
``tta_ensemble.py:94-110`` (``_add_full_channels``) with stock torch ops.  Pinned by ``tests/golden/tta_goldens.npz``:
``oracle/make_tta_goldens.py`` runs the REAL ``tta_combinations.py`` and ``tta_ensemble.py`` in place (``ref_loader``) and
stores their augmentation lists, mode maps and ensemble results; ``tests/test_oracle_tta.py`` checks this file and the
product's host logic against them."""

from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch


def view(x: torch.Tensor, flip_axes, rotation_plane, k: int) -> torch.Tensor:
    """tta.py:706-714 — spatial axes 0..2 map to tensor dims 2..4."""
    if flip_axes:
        x = torch.flip(x, dims=[int(a) + 2 for a in flip_axes])
    if rotation_plane is not None and k > 0:
        x = torch.rot90(x, k=k, dims=(int(rotation_plane[0]) + 2, int(rotation_plane[1]) + 2))
    return x


def invert_view(pred: torch.Tensor, flip_axes, rotation_plane, k: int) -> torch.Tensor:
    """tta_affinity.py:364-369."""
    if rotation_plane is not None and int(k) % 4:
        pred = torch.rot90(pred, k=-int(k), dims=(int(rotation_plane[0]) + 2, int(rotation_plane[1]) + 2))
    if flip_axes:
        pred = torch.flip(pred, dims=[int(a) + 2 for a in flip_axes])
    return pred


def apply_preprocessing(t: torch.Tensor, act_codes: Sequence[int], act_scales: Sequence[float],
                        select: Optional[Sequence[int]], output_dtype: torch.dtype) -> torch.Tensor:
    """tta.py:312-402 with the activations given as per-channel codes (0 none, 1 sigmoid, 2 scale_sigmoid, 3 tanh)."""
    t = t.clone()
    for c, (code, scale) in enumerate(zip(act_codes, act_scales)):
        ch = t[:, c:c + 1]
        if code == 1:
            ch.sigmoid_()
        elif code == 2:
            ch.mul_(scale).sigmoid_()
        elif code == 3:
            ch.tanh_()
    if select is not None:
        t = t[:, list(select)]
    return t if t.dtype == output_dtype else t.to(output_dtype)


def fold(current: Optional[torch.Tensor], incoming: torch.Tensor, modes: Sequence[str], n_prev: int) -> torch.Tensor:
    """tta_ensemble.py:94-110 (non-distributed)."""
    if current is None or n_prev == 0:
        return incoming.clone()
    for c, mode in enumerate(modes):
        cur, inc = current[:, c:c + 1], incoming[:, c:c + 1]
        if mode == "mean":
            delta = inc - cur
            cur += delta / (n_prev + 1)
        elif mode == "min":
            cur.copy_(torch.minimum(cur, inc))
        elif mode == "max":
            cur.copy_(torch.maximum(cur, inc))
        else:
            raise ValueError(f"Unknown TTA ensemble modes: ['{mode}'].")
    return current


def tta_predict(images: torch.Tensor, network_fn: Callable, combos, modes: List[str], act_codes, act_scales,
                select, output_dtype: torch.dtype) -> torch.Tensor:
    acc = None
    for n_prev, (flip_axes, plane, k) in enumerate(combos):
        pred = network_fn(view(images, flip_axes, plane, k))
        pred = invert_view(pred, flip_axes, plane, k)
        acc = fold(acc, apply_preprocessing(pred, act_codes, act_scales, select, output_dtype), modes, n_prev)
    return acc


# ----------------------------------------------------------------------------- affinity-aware path (round 2)
def ramp_network(num_channels: int):
    """Deterministic stand-in for the model in the TTA tests and goldens: ``out[:, c] = x[:, 0] * ramp_c + 0.1 c`` with a
    position-dependent ramp built from the VIEW tensor's own shape, so that a wrong inverse view or channel move changes
    the result.  Only fp32 multiplies and adds (no FMA in eager torch): bit-identical on CPU and CUDA."""

    def fn(x: torch.Tensor) -> torch.Tensor:
        d, h, w = (int(v) for v in x.shape[2:])
        dev = x.device
        z = torch.arange(d, device=dev, dtype=torch.float32).view(d, 1, 1)
        y = torch.arange(h, device=dev, dtype=torch.float32).view(1, h, 1)
        xx = torch.arange(w, device=dev, dtype=torch.float32).view(1, 1, w)
        outs = []
        for c in range(num_channels):
            ramp = z * (0.03125 * (c + 1)) + y * 0.0625 - xx * (0.015625 * (c + 2)) + 0.5
            outs.append(x[:, 0].float() * ramp + 0.125 * c)
        return torch.stack(outs, dim=1)

    return fn


def preprocess_specs(t: torch.Tensor, specs, select, output_dtype: torch.dtype) -> torch.Tensor:
    """tta.py:327-402 for ``specs = [(channel list, activation name), ...]`` applied one after the other in place."""
    t = t.clone()
    for chans, act in specs:
        chans = list(chans)
        if act == "sigmoid":
            t[:, chans] = torch.sigmoid(t[:, chans])
        elif act == "tanh":
            t[:, chans] = torch.tanh(t[:, chans])
        elif isinstance(act, str) and act.startswith("scale_sigmoid"):
            scale = float(act.split(":", 1)[1]) if ":" in act else 0.2
            t[:, chans] = torch.sigmoid(scale * t[:, chans])
        elif act == "softmax" and len(chans) > 1:
            t[:, chans] = torch.softmax(t[:, chans], dim=1)
    if select is not None:
        t = t[:, list(select)]
    return t if t.dtype == output_dtype else t.to(output_dtype)
