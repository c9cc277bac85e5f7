"""Fused AdamW + gradient-norm clip + EMA over flat arenas.

The reference builds ``torch.optim.AdamW`` with one param group per parameter (``training/optimization/build.py:47-113``:
norm-layer parameters use ``weight_decay_norm`` (default 0), biases ``weight_decay_bias`` and ``lr * bias_lr_factor``),
clips with Lightning's ``gradient_clip_val`` (tutorials: 1.0) and keeps an EMA of the weights in a callback
(``training/lightning/callbacks.py:869-907``) — three full passes over parameters and optimizer state plus a norm
reduction, as ~1 000 small launches.  Here parameters, gradients (``FlatGradArena``), both moments and the EMA each live
in ONE contiguous fp32 arena and a step is two launches (``pcb_grad_sumsq`` when clipping, ``pcb_adamw_step``); the
1/world gradient scale of the DDP mean is folded in, the step counter lives on the device (CUDA-graph capturable).
"""

from __future__ import annotations

import ctypes
from typing import Dict, Iterable, List, Optional

import torch

from .. import _lib as L
from .ddp import FlatGradArena

_NORM_TYPES = (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d, torch.nn.BatchNorm3d, torch.nn.SyncBatchNorm, torch.nn.GroupNorm,
               torch.nn.InstanceNorm1d, torch.nn.InstanceNorm2d, torch.nn.InstanceNorm3d, torch.nn.LayerNorm,
               torch.nn.LocalResponseNorm)


def reference_param_groups(model: torch.nn.Module, lr: float, weight_decay: float, weight_decay_norm: float = 0.0,
                           weight_decay_bias: Optional[float] = None, bias_lr_factor: float = 1.0) -> List[dict]:
    """``build.py:69-111``: one group per parameter, in ``model.modules()`` order, shared parameters once."""
    wdb = weight_decay if weight_decay_bias is None else weight_decay_bias
    groups, memo = [], set()
    for module in model.modules():
        # exactly the reference's test: torch's own norm classes.  The channels-first ``LayerNorm`` of MedNeXt's
        # ``norm_type="layer"`` is a plain ``nn.Module`` upstream, so its weight is decayed like any other weight and its bias
        # takes the bias rule — reproduced here, checked against the real ``build_optimizer`` in tests/test_host_logic.py
        norm = isinstance(module, _NORM_TYPES)
        for key, value in module.named_parameters(recurse=False):
            if not value.requires_grad or value in memo:
                continue
            memo.add(value)
            g_lr, g_wd = lr, weight_decay
            if norm:
                g_wd = weight_decay_norm
            elif key == "bias":
                g_lr, g_wd = lr * bias_lr_factor, wdb
            groups.append({"params": [value], "lr": g_lr, "weight_decay": g_wd})
    return groups


class FusedAdamW:
    """AdamW over flat arenas.  ``groups``: list of ``{"params": [...], "lr": ..., "weight_decay": ...}`` (torch format).
    The parameters' storage is MOVED into one flat arena (``p.data`` becomes a view of it), so the modules keep working
    unchanged; ``arena`` is the gradient arena the backward pass writes into (created over the same parameter order when
    not given)."""

    def __init__(self, groups: Iterable[dict], *, betas=(0.9, 0.999), eps: float = 1e-8, max_grad_norm: float = 0.0,
                 ema_decay: Optional[float] = None, arena: Optional[FlatGradArena] = None, world_size: int = 1,
                 ema_warmup_steps: int = 0):
        groups = [dict(g) for g in groups]
        params: List[torch.nn.Parameter] = [p for g in groups for p in g["params"] if p.requires_grad]
        if not params:
            raise ValueError("FusedAdamW: no trainable parameters")
        dev = params[0].device
        L.require_device(params[0], "FusedAdamW")
        if any(p.dtype != torch.float32 or p.device != dev for p in params):
            raise ValueError("FusedAdamW expects fp32 parameters on one CUDA device")
        self.params = params
        if arena is not None and sorted(id(p) for p in arena.params) != sorted(id(p) for p in params):
            raise ValueError("FusedAdamW: the gradient arena must cover the same parameters")
        self.arena = arena if arena is not None else FlatGradArena(params)
        if self.arena.align % 4:
            raise ValueError("FusedAdamW needs a gradient arena with 16-byte aligned slices (align % 4 == 0)")
        n = self.arena.total                        # padded layout of the gradient arena (16-byte aligned slices)
        self.n = n
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        by_id = {id(p): (g["lr"], g["weight_decay"]) for g in groups for p in g["params"]}
        ends, lrs, wds = [], [], []
        for i, p in enumerate(self.arena.params):
            view = self.arena.view_of(i, self.flat)
            view.copy_(p.data)
            p.data = view
            nxt = self.arena.offsets[i + 1] if i + 1 < len(self.arena.params) else n
            ends.append(nxt)                           # the padding behind a parameter belongs to its segment (stays 0)
            lrs.append(float(by_id[id(p)][0])); wds.append(float(by_id[id(p)][1]))
        self.seg_end = torch.tensor(ends, device=dev, dtype=torch.int64)
        self.seg_lr = torch.tensor(lrs, device=dev, dtype=torch.float32)
        self.seg_wd = torch.tensor(wds, device=dev, dtype=torch.float32)
        self._base_lr = list(lrs)
        self.exp_avg = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(n, device=dev, dtype=torch.float32)
        self.ema = self.flat.clone() if ema_decay is not None else None
        self.ema_decay = float(ema_decay) if ema_decay is not None else 0.0
        # callbacks.py:815-817: the first `warmup_steps` updates use decay 0 (the EMA tracks the weights exactly).  The decay
        # is a launch argument, so the count lives on the host: capture a CUDA graph of the step AFTER the warm-up.
        self.ema_warmup_steps = max(0, int(ema_warmup_steps))
        self.ema_updates = 0
        self.step_count = torch.zeros(1, device=dev, dtype=torch.float32)
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float64)
        self.seg_active = torch.ones(len(ends), device=dev, dtype=torch.int32)      # device scratch of the kernel
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.max_grad_norm = float(max_grad_norm or 0.0)
        self.world_size = int(world_size)

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.arena.zero_grad(set_to_none)

    def set_lr_scale(self, scale: float) -> None:
        """learning-rate schedule hook: every segment's lr = its base lr x ``scale`` (one tiny H2D copy)."""
        self.seg_lr.copy_(torch.tensor([v * float(scale) for v in self._base_lr], dtype=torch.float32), non_blocking=True)

    @torch.no_grad()
    def step(self, grads_are_summed: bool = False) -> None:
        """One optimizer step on the current stream.  ``grads_are_summed``: the arena holds the all-reduced SUM over
        ``world_size`` ranks (``arena.allreduce_sum``), the 1/world of the DDP mean is folded into the kernel."""
        lib, st = L.lib(), L.stream_ptr(self.flat.device)
        scale = 1.0 / self.world_size if grads_are_summed and self.world_size > 1 else 1.0
        clip = self.max_grad_norm > 0.0
        self.ema_updates += 1
        ema_decay = 0.0 if self.ema_updates <= self.ema_warmup_steps else self.ema_decay
        if clip:
            self.sumsq.zero_()
            L.check(lib.pcb_grad_sumsq(L.ptr(self.arena.buffer), ctypes.c_int64(self.n), L.ptr(self.sumsq), st), "pcb_grad_sumsq")
        L.check(lib.pcb_adamw_step(L.ptr(self.flat), L.ptr(self.arena.buffer), L.ptr(self.exp_avg), L.ptr(self.exp_avg_sq),
                                   L.ptr(self.ema), ctypes.c_int64(self.n), L.ptr(self.seg_end), L.ptr(self.seg_lr),
                                   L.ptr(self.seg_wd), ctypes.c_int(int(self.seg_end.numel())), ctypes.c_float(self.betas[0]),
                                   ctypes.c_float(self.betas[1]), ctypes.c_float(self.eps), L.ptr(self.step_count),
                                   L.ptr(self.sumsq) if clip else None, ctypes.c_float(self.max_grad_norm),
                                   ctypes.c_float(scale), ctypes.c_float(ema_decay), L.ptr(self.seg_active), st),
                "pcb_adamw_step")
        L.PARAM_EPOCH[0] += 1          # parameters changed behind autograd's back: kernel-layout weight caches repack

    def grad_norm(self) -> torch.Tensor:
        """the total norm the last clipped step saw (device scalar)"""
        return self.sumsq.sqrt()

    def ema_tensors(self) -> Dict[int, torch.Tensor]:
        """``{id(param): ema view}`` (what the EMA callback swaps in for validation)"""
        if self.ema is None:
            return {}
        return {id(p): self.arena.view_of(i, self.ema) for i, p in enumerate(self.arena.params)}

    def swap_ema(self) -> None:
        """exchange live weights and EMA weights in place (``callbacks.py:909-944``: validate with EMA, then swap back)"""
        if self.ema is None:
            raise RuntimeError("FusedAdamW was built without ema_decay")
        tmp = self.flat.clone()
        self.flat.copy_(self.ema)
        self.ema.copy_(tmp)
        L.PARAM_EPOCH[0] += 1


def build_fused_adamw(cfg, model: torch.nn.Module, *, arena: Optional[FlatGradArena] = None, world_size: int = 1,
                      ema_decay: Optional[float] = None, ema_warmup_steps: int = 0) -> FusedAdamW:
    """``build_optimizer(cfg, model)`` (``build.py:47-113``) for ``optimizer.name == 'adamw'`` on the fused kernel."""
    if not (hasattr(cfg, "optimization") and hasattr(cfg.optimization, "optimizer")):
        raise ValueError("Config must have 'optimization.optimizer' section")
    oc = cfg.optimization.optimizer
    name = str(getattr(oc, "name", "adamw")).lower()
    if name != "adamw":
        raise NotImplementedError(f"pcb200 fused optimizer implements 'adamw' only (got {name!r}); use torch.optim for others")
    lr = float(getattr(oc, "lr", 1e-4))
    wd = float(getattr(oc, "weight_decay", 1e-4))
    groups = reference_param_groups(model, lr, wd, float(getattr(oc, "weight_decay_norm", 0.0)),
                                    getattr(oc, "weight_decay_bias", None), float(getattr(oc, "bias_lr_factor", 1.0)))
    if arena is not None:      # keep the arena's parameter order
        order = {id(p): i for i, p in enumerate(arena.params)}
        groups.sort(key=lambda g: order[id(g["params"][0])])
    clip = float(getattr(cfg.optimization, "gradient_clip_val", 0.0) or 0.0)      # trainer.py:321
    return FusedAdamW(groups, betas=tuple(getattr(oc, "betas", (0.9, 0.999))), eps=float(getattr(oc, "eps", 1e-8)),
                      max_grad_norm=clip, ema_decay=ema_decay, arena=arena, world_size=world_size,
                      ema_warmup_steps=ema_warmup_steps)
