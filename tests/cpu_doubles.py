"""TEST INFRASTRUCTURE — CPU stand-ins for the handful of kernel-calling helpers of ``inference/window.py`` (crop, weight map,
accumulate, normalise) built from the oracle's restatements of the same reference lines, plus the oracle ensemble for
``TTAEnsemble.predict``.  They let the HOST logic that sits above those helpers (the cfg-driven lazy seam, the chunked driver,
the TTA predictor) run end to end in the CPU suite.  They are never the thing under test and never reachable from the product
package: the GPU tests run the same calls against the real kernels."""

from __future__ import annotations

import torch

from oracle import window_oracle as O


def _extract_starts(tensor, starts, roi, padding_mode, cval):
    return torch.cat([O.extract_patch(tensor, tuple(int(v) for v in s), tuple(roi), padding_mode, cval) for s in starts], dim=0)


def _accumulate_window(pred, wmap, value, weight, roi, out_size, pred_lo, out_lo, box):
    """``value[:, out box] += pred[box] * map[box]; weight[out box] += map[box]`` — multiply, then add (window.py:648-655)"""
    psl = tuple(slice(int(l), int(l) + int(b)) for l, b in zip(pred_lo, box))
    osl = tuple(slice(int(l), int(l) + int(b)) for l, b in zip(out_lo, box))
    w = wmap[psl]
    value[(0, slice(None)) + osl] += pred[(slice(None),) + psl] * w
    weight[(0, 0) + osl] += w


def _accumulate_batch(pred, wmap, value, weight, roi, out_size, starts):
    for i, s in enumerate(starts):
        _accumulate_window(pred[i], wmap, value, weight, roi, out_size, (0,) * len(roi), s, roi)


def build_sliding_importance_map(roi_size, *, mode, device=None, dtype=torch.float32, min_value=1e-5):
    from pytorch_connectomics_b200.inference import window as W
    m = W._normalize_blending_mode(mode)
    return O.importance_map(roi_size, "distance_transform" if m in W._DISTANCE_TRANSFORM_BLEND_MODES else m, dtype,
                            0.0 if m in W._DISTANCE_TRANSFORM_BLEND_MODES else min_value)


def install(monkeypatch) -> None:
    """route the window helpers (and ``TTAEnsemble.predict``) to the CPU stand-ins for the duration of one test"""
    from pytorch_connectomics_b200.inference import tta as T
    from pytorch_connectomics_b200.inference import window as W
    monkeypatch.setattr(W, "_device_or_raise", lambda device: torch.device(device))
    monkeypatch.setattr(W, "_extract_starts", _extract_starts)
    monkeypatch.setattr(W, "_accumulate_window", _accumulate_window)
    monkeypatch.setattr(W, "_accumulate_batch", _accumulate_batch)
    monkeypatch.setattr(W, "build_sliding_importance_map", build_sliding_importance_map)
    monkeypatch.setattr(W, "normalize_weighted_accumulator", lambda v, w: O.normalize_accumulator(v, w))
    monkeypatch.setattr(T.TTAEnsemble, "predict", oracle_ensemble_predict)


def oracle_ensemble_predict(self, images, network_fn):
    """``TTAEnsemble.predict`` with the fold kernels replaced by the oracle chain of ``oracle/tta_oracle.py`` (view ->
    network -> inverse view -> activations incl. softmax groups -> channel selection -> output dtype -> running mean / min /
    max); the configuration resolvers are the product's own."""
    from oracle import tta_oracle as TO
    from pytorch_connectomics_b200.inference.tta import (_resolve_ensemble_mode_map, resolve_activation_specs,
                                                         resolve_channel_indices)
    combos = self.combinations(images.dim())
    acc, modes = None, None
    for n_prev, (flip_axes, plane, k) in enumerate(combos):
        pred = TO.invert_view(network_fn(TO.view(images, flip_axes, plane, k)), flip_axes, plane, k)
        n_raw = int(pred.shape[1])
        codes, scales, groups = resolve_activation_specs(self.channel_activations, n_raw)
        sel = resolve_channel_indices(self.select_channel, num_channels=n_raw, context="inference.model.select_channel")
        t = pred.clone()
        done = set()
        for c in range(n_raw):
            if codes[c] == 4 and tuple(groups[c]) not in done:          # softmax over the spec's channel list, once
                done.add(tuple(groups[c]))
                t[:, groups[c]] = torch.softmax(pred[:, groups[c]], dim=1)
        t = TO.apply_preprocessing(t, [0 if c == 4 else c for c in codes], scales, sel, self.output_dtype or torch.float32)
        if modes is None:
            mode_cfg = getattr(self.tta_cfg, "ensemble_mode", "mean") if self.tta_cfg is not None else "mean"
            modes = _resolve_ensemble_mode_map(mode_cfg, int(t.shape[1]))
        acc = TO.fold(acc, t, modes, n_prev)
    return acc
