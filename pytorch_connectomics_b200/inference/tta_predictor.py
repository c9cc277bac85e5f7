"""``TTAPredictor(cfg, sliding_inferer, forward_fn)`` — the object the reference's inference loop and lazy path construct and
call (``connectomics/inference/tta.py:67-79`` constructor, ``:1619-1666`` ``predict``), over the B200 fold kernels.

It reads the same config nodes (``inference.test_time_augmentation.{enabled, flip_axes, rotation90_axes, rotate90_k,
ensemble_mode, patch_first_local, distributed_sharding, apply_mask}``, ``inference.model.{channel_activations,
select_channel, output_dtype, head}``, ``model.primary_head``) and keeps the call order of ``_predict_prepared_tensor``
(``:806-878``) / ``_predict_patch_first_local`` (``:880-1314``): network (through the sliding-window engine when there is
one) per view -> inverse view -> activations -> channel selection -> output dtype -> ensemble -> mask.  Everything between
"network" and "mask" is :class:`~.tta.TTAEnsemble` (one fold kernel per view); with TTA disabled the same fold runs once on
the identity view, which is ``apply_preprocessing`` (``:312-402``).  The mask step (``:465-548,1568-1617``) is restated here:
shape checks with the reference's messages, optional centre pad/crop, ``tanh`` channels are filled with -1 outside the mask.
"""

from __future__ import annotations

import logging
from typing import Any, Callable, List, Optional

import numpy as np
import torch

from . import window as W
from .model_outputs import pick_inference_output, resolve_output_head
from .tta import TTAEnsemble, resolve_activation_specs, resolve_channel_indices

logger = logging.getLogger(__name__)

_ACT_NAMES = {0: None, 1: "sigmoid", 2: "scale_sigmoid", 3: "tanh", 4: "softmax"}


def _node(obj: Any, *path: str, default: Any = None) -> Any:
    for key in path:
        if obj is None:
            return default
        obj = obj.get(key, None) if isinstance(obj, dict) else getattr(obj, key, None)
    return default if obj is None else obj


class _EngineNetwork:
    """What the predictor hands to the sliding-window engine as ``network``: calls ``_sliding_window_predict`` and, when
    ``forward_fn`` is a pcb200 MedNeXt whose single output needs no head selection, exposes its ``native_plan`` so the
    engine takes the in-library tile loop (``pcb_sw_run``) instead of the generic one (``window.native_plan_of``)."""

    def __init__(self, predictor: "TTAPredictor") -> None:
        self._p = predictor

    def __call__(self, inputs: torch.Tensor) -> torch.Tensor:
        return self._p._sliding_window_predict(inputs)

    def native_plan(self):
        p = self._p
        if p._requested_output_head_override is not None:
            return None
        # forward_fn is the module itself, its bound ``forward``, or the bound ``forward`` of the reference's LightningModule
        # (``training/lightning/model.py:236-242``: a pure ``return self.model(x)``) — nothing else is looked through
        fn = p.forward_fn
        owners = [fn]
        bound = getattr(fn, "__self__", None)
        if bound is not None and getattr(fn, "__name__", "") == "forward":
            owners.append(bound)
            if type(bound).__name__ == "ConnectomicsModule":
                owners.append(getattr(bound, "model", None))
        getter = next((g for g in (getattr(o, "native_plan", None) for o in owners if o is not None) if callable(g)), None)
        if getter is None:
            return None
        plan = getter()
        if plan is not None and len(getattr(plan, "head_channels", ())) == 1 and p._seen_raw is None:
            p._seen_raw = int(plan.head_channels[0])
            p._record_activation_types(p._seen_raw)
        return plan


class TTAPredictor:
    """Same constructor and ``predict`` signature as the reference's class."""

    def __init__(self, cfg, sliding_inferer, forward_fn: Callable[[torch.Tensor], Any]) -> None:
        self.cfg = cfg
        self.sliding_inferer = sliding_inferer
        self.forward_fn = forward_fn
        self.channel_activation_types: Optional[List[Optional[str]]] = None
        self._requested_output_head_override: Optional[str] = None
        self._last_distributed_sharding_active = False
        self._last_skip_postprocess_on_rank = False
        self._seen_raw: Optional[int] = None
        self._engine_network = _EngineNetwork(self)

    # ---- config ------------------------------------------------------------------------------------------------------
    def _get_tta_cfg(self):
        return _node(self.cfg, "inference", "test_time_augmentation")

    @staticmethod
    def _distributed_context():
        from .lazy_distributed import distributed_context
        return distributed_context()

    def is_distributed_sharding_enabled(self) -> bool:
        """``tta.py:246-256``"""
        tta = self._get_tta_cfg()
        active, _, world = self._distributed_context()
        return bool(tta is not None and getattr(tta, "enabled", False) and getattr(tta, "distributed_sharding", False)
                    and active and world > 1)

    def should_skip_postprocess_on_rank(self) -> bool:
        """``tta.py:258-260`` — True on non-root ranks after a sharded ensemble was reduced onto rank 0."""
        return self._last_distributed_sharding_active and self._last_skip_postprocess_on_rank

    def _is_patch_first_local_tta_enabled(self) -> bool:
        tta = self._get_tta_cfg()
        return bool(tta is not None and getattr(tta, "enabled", False) and getattr(tta, "patch_first_local", False)
                    and self.sliding_inferer is not None)

    def _output_dtype(self) -> Optional[torch.dtype]:
        return W.resolve_model_output_dtype(self.cfg) if _node(self.cfg, "inference") is not None else None

    def _ensemble(self, tta_cfg, *, sharded: bool) -> TTAEnsemble:
        return TTAEnsemble(tta_cfg, channel_activations=_node(self.cfg, "inference", "model", "channel_activations"),
                           select_channel=_node(self.cfg, "inference", "model", "select_channel"),
                           output_dtype=self._output_dtype(), cfg=self.cfg, requested_head=self._requested_output_head_override,
                           distributed_sharding=sharded)

    def _record_activation_types(self, num_raw: int) -> None:
        """per OUTPUT channel activation names after selection (``tta.py:383-399``) — what the mask step keys on"""
        acts = _node(self.cfg, "inference", "model", "channel_activations")
        if not acts:
            self.channel_activation_types = None
            return
        codes, _scales, _groups = resolve_activation_specs(acts, num_raw)
        names: List[Optional[str]] = [_ACT_NAMES.get(int(c)) for c in codes]
        sel = resolve_channel_indices(_node(self.cfg, "inference", "model", "select_channel"), num_channels=num_raw,
                                      context="inference.model.select_channel")
        if sel is not None:
            names = [names[i] for i in sel]
        self.channel_activation_types = names if any(n is not None for n in names) else None

    # ---- network -----------------------------------------------------------------------------------------------------
    def _sliding_window_predict(self, inputs: torch.Tensor) -> torch.Tensor:
        """``tta.py:449-463`` — one forward, then the requested / primary head's tensor."""
        with torch.no_grad():
            out = self.forward_fn(inputs)
        pred = pick_inference_output(self.cfg, out, self._requested_output_head_override)
        num_raw = int(pred.shape[1])
        if self._seen_raw != num_raw:
            self._seen_raw = num_raw
            self._record_activation_types(num_raw)
        return pred

    def _run_network(self, images: torch.Tensor) -> torch.Tensor:
        """``tta.py:415-433``"""
        if self.sliding_inferer is not None:
            return self.sliding_inferer(inputs=images, network=self._engine_network)
        if bool(_node(self.cfg, "inference", "sliding_window", "keep_input_on_cpu", default=False)) and images.device.type == "cpu":
            raise RuntimeError("inference.sliding_window.keep_input_on_cpu=True requires sliding-window inference to be "
                               "enabled (set inference.sliding_window.window_size or model output size).")
        return self._sliding_window_predict(images)

    # ---- input / mask ------------------------------------------------------------------------------------------------
    def _normalize_input(self, images: torch.Tensor) -> torch.Tensor:
        """``tta.py:580-601``"""
        if images.ndim == 3:
            images = images[None, None]
        elif images.ndim == 4:
            images = images[:, None]
        elif images.ndim != 5:
            raise ValueError(f"TTA requires 3D, 4D, or 5D input tensor. Got {images.ndim}D tensor with shape {images.shape}. "
                             "Expected shapes: (D, H, W), (B, D, H, W), or (B, C, D, H, W)")
        if W.is_2d_inference_mode(self.cfg) and images.size(2) == 1:
            raise NotImplementedError("pcb200 TTAPredictor: 2-D inference mode is not implemented in the B200 engine (3-D only)")
        return images

    def _coerce_mask_to_tensor(self, mask: Any) -> torch.Tensor:
        """``tta.py:550-574`` — unwrap what dataloader collation nests around a mask volume"""
        while isinstance(mask, (list, tuple)) and len(mask) == 1:
            mask = mask[0]
        if isinstance(mask, np.ndarray):
            return torch.from_numpy(mask)
        if torch.is_tensor(mask):
            return mask
        if isinstance(mask, (list, tuple)):
            parts = [self._coerce_mask_to_tensor(m) for m in mask]
            if not parts:
                raise ValueError("Mask list is empty after collation.")
            try:
                return torch.stack(parts)
            except RuntimeError:
                raise ValueError("Mask list contains tensors with incompatible shapes for stacking: "
                                 f"{[tuple(t.shape) for t in parts]}") from None
        raise TypeError(f"Unsupported mask type: {type(mask).__name__}")

    def _validate_and_prepare_mask(self, mask, prediction: torch.Tensor, align_to_image: bool = False) -> torch.Tensor:
        """``tta.py:465-548`` — rank / batch / channel checks, strict spatial match unless ``align_to_image`` (centre crop or
        zero pad per axis), binarised in the prediction's dtype."""
        if mask is None:
            raise ValueError("Mask is None while mask application is enabled.")
        mask = self._coerce_mask_to_tensor(mask).to(prediction.device, non_blocking=True)
        if mask.ndim == prediction.ndim - 1:
            mask = mask.unsqueeze(1)
        elif mask.ndim == prediction.ndim - 2:
            mask = mask[None, None]
        if mask.ndim != prediction.ndim:
            raise ValueError(f"Mask rank {mask.ndim} does not match prediction rank {prediction.ndim}. "
                             f"mask.shape={tuple(mask.shape)}, prediction.shape={tuple(prediction.shape)}")
        if mask.shape[0] != prediction.shape[0]:
            if mask.shape[0] != 1:
                raise ValueError(f"Mask batch {mask.shape[0]} does not match prediction batch {prediction.shape[0]}.")
            mask = mask.expand(prediction.shape[0], *mask.shape[1:])
        if mask.shape[1] not in (1, prediction.shape[1]):
            raise ValueError(f"Mask channels {mask.shape[1]} incompatible with prediction channels "
                             f"{prediction.shape[1]}. Expected C=1 or C={prediction.shape[1]}.")
        if mask.shape[2:] != prediction.shape[2:]:
            if not align_to_image:
                raise ValueError("Mask spatial shape must exactly match prediction spatial shape. "
                                 f"Got mask.shape={tuple(mask.shape)} and prediction.shape={tuple(prediction.shape)}. "
                                 "Fix test/tune mask preprocessing so they produce identical spatial dimensions.")
            nsp = mask.ndim - 2
            for ax in range(nsp):
                have, want = int(mask.shape[2 + ax]), int(prediction.shape[2 + ax])
                if have > want:                          # centre crop
                    lo = (have - want) // 2
                    mask = mask.narrow(2 + ax, lo, want)
                elif have < want:                        # zero pad, the odd voxel goes behind
                    before = (want - have) // 2
                    pad = [0, 0] * nsp
                    pad[2 * (nsp - 1 - ax)], pad[2 * (nsp - 1 - ax) + 1] = before, want - have - before
                    mask = torch.nn.functional.pad(mask, tuple(pad), mode="constant", value=0)
        return (mask > 0).to(dtype=prediction.dtype)

    def _apply_mask_to_result(self, result: torch.Tensor, mask, mask_align_to_image: bool) -> torch.Tensor:
        """``tta.py:1568-1617`` — multiply by the mask; channels that went through ``tanh`` are set to -1 outside it."""
        tta = self._get_tta_cfg()
        if mask is None or not (getattr(tta, "apply_mask", True) if tta is not None else True):
            return result
        try:
            m = self._validate_and_prepare_mask(mask, result, align_to_image=mask_align_to_image)
        except TypeError as exc:
            logger.warning("Skipping mask application because the provided mask payload is not a tensor-like volume: %s", exc)
            return result
        types = self.channel_activation_types
        if types is None or len(types) != int(result.shape[1]):
            return result * m
        for c, act in enumerate(types):
            mc = m[:, c:c + 1] if m.shape[1] == result.shape[1] else m[:, 0:1]
            if act == "tanh":
                result[:, c:c + 1] = mc * result[:, c:c + 1] + (1 - mc) * (-1.0)
            else:
                result[:, c:c + 1] = mc * result[:, c:c + 1]
        return result

    # ---- the public calls --------------------------------------------------------------------------------------------
    def apply_preprocessing(self, tensor: torch.Tensor) -> torch.Tensor:
        """``tta.py:312-402`` — activations, channel selection and output dtype of ONE prediction (the identity view through
        the fold kernel)."""
        if _node(self.cfg, "inference") is None:
            return tensor
        self._record_activation_types(int(tensor.shape[1]))
        return self._ensemble(None, sharded=False).predict(tensor, lambda t: t)

    def predict(self, images: torch.Tensor, mask=None, mask_align_to_image: bool = False,
                requested_head: Optional[str] = None) -> torch.Tensor:
        """``tta.py:1619-1666``"""
        previous = self._requested_output_head_override
        if requested_head is not None:               # tta.py:1640-1646: an explicit request must name a configured head
            resolve_output_head(self.cfg, requested_head=requested_head, purpose="inference output selection", allow_none=False)
        self._requested_output_head_override = requested_head
        self._seen_raw = None
        try:
            images = self._normalize_input(images)
            self._last_distributed_sharding_active = False
            self._last_skip_postprocess_on_rank = False
            tta = self._get_tta_cfg()
            enabled = tta is not None and getattr(tta, "enabled", True)
            if not enabled:
                result = self._ensemble(None, sharded=False).predict(images, self._run_network)
                return self._apply_mask_to_result(result, mask, mask_align_to_image)
            sharded = self.is_distributed_sharding_enabled()
            ens = self._ensemble(tta, sharded=sharded)
            combos = ens.combinations(images.dim())
            single = len(combos) == 1 and not list(combos[0][0]) and (combos[0][1] is None or int(combos[0][2]) % 4 == 0)
            self._last_distributed_sharding_active = sharded and not single
            if single:
                ens = self._ensemble(tta, sharded=False)
            if self._is_patch_first_local_tta_enabled() and not single:
                inf = self.sliding_inferer
                result = ens.predict_patch_first(images, self._sliding_window_predict, roi_size=inf.roi_size,
                                                 overlap=inf.overlap, sw_batch_size=inf.sw_batch_size, mode=inf.mode,
                                                 padding_mode=inf.padding_mode, cval=inf.cval)
            else:
                result = ens.predict(images, self._run_network)
            if self._last_distributed_sharding_active and result.numel() == 0:
                self._last_skip_postprocess_on_rank = True
                return result
            return self._apply_mask_to_result(result, mask, mask_align_to_image)
        finally:
            self._requested_output_head_override = previous


__all__ = ["TTAPredictor"]
