"""The optimizer-step semantics of the reference's Lightning loop on the flat arenas (``training/lightning/trainer.py:231-256,
314-334``: ``DDPStrategy`` + ``accumulate_grad_batches`` + ``gradient_clip_val``; ``lightning/model.py:863-910`` training_step).

Lightning's automatic optimisation divides the loss by ``accumulate_grad_batches``, lets DDP synchronise gradients only on the LAST
micro-batch of a window (``no_sync`` before), clips the total gradient norm and steps.  Here the window accumulates into the flat
gradient arena (autograd adds in place), the one exchange is ONE NCCL all-reduce (SUM) of the arena after the last micro-batch,
and the 1/world of the DDP mean, the norm clip, AdamW and the EMA update are one kernel (``FusedAdamW.step``)."""

from __future__ import annotations

from typing import Callable, Optional

import torch

from .data_parallel import ArenaDataParallel
from .ddp import FlatGradArena
from .optim import FusedAdamW


class ArenaTrainStep:
    """``step(x, target) -> loss`` for one MICRO-batch; the optimizer runs every ``accumulate_grad_batches`` calls.

    ``loss_fn(model(x), target)`` is any differentiable scalar.  ``group``: the data-parallel process group (None = default;
    without an initialised group the step is single-process).

    ``model`` may be an :class:`~.data_parallel.ArenaDataParallel` built on the optimizer's arena: then the exchange is the
    wrapper's (segments all-reduced while the last backward of the window still runs, nothing inside ``no_sync()`` on the
    other micro-batches) and this class only steps; the wrapper's ``reduce_op`` decides whether the 1/world is already in the
    gradients (``"mean"``) or rides in the optimizer kernel (``"sum"``)."""

    def __init__(self, model: torch.nn.Module, loss_fn: Callable, optimizer: FusedAdamW, *, accumulate_grad_batches: int = 1,
                 group=None) -> None:
        if int(accumulate_grad_batches) < 1:
            raise ValueError(f"accumulate_grad_batches must be >= 1, got {accumulate_grad_batches}")
        if not isinstance(optimizer, FusedAdamW):
            raise TypeError("ArenaTrainStep drives a FusedAdamW (build_fused_adamw); use GraphedTrainStep for torch optimizers")
        self.model, self.loss_fn, self.opt = model, loss_fn, optimizer
        self.arena: FlatGradArena = optimizer.arena
        self.wrapper: Optional[ArenaDataParallel] = model if isinstance(model, ArenaDataParallel) else None
        if self.wrapper is not None and self.wrapper.arena is not self.arena:
            raise ValueError("ArenaTrainStep: build the ArenaDataParallel on the optimizer's gradient arena (arena=optimizer.arena)")
        self.k = int(accumulate_grad_batches)
        self.group = group
        self.micro = 0
        self.optimizer_steps = 0

    @property
    def will_step(self) -> bool:
        """True when the next call closes an accumulation window."""
        return self.micro == self.k - 1

    def __call__(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if self.micro == 0:
            self.arena.zero()
        boundary = self.micro == self.k - 1
        if self.wrapper is not None and not boundary:
            with self.wrapper.no_sync():
                loss = self.loss_fn(self.model(x), target)
                (loss / self.k if self.k > 1 else loss).backward()
        else:
            loss = self.loss_fn(self.model(x), target)
            (loss / self.k if self.k > 1 else loss).backward()
        self.micro += 1
        if self.micro == self.k:
            self.micro = 0
            self._exchange_and_step(exchanged=self.wrapper is not None)
        return loss.detach()

    def _exchange_and_step(self, exchanged: bool) -> None:
        if self.wrapper is not None:
            if not exchanged:
                self.wrapper.reduce_now()
            self.opt.step(grads_are_summed=self.wrapper.reduce_op == "sum")
        else:
            self.arena.gather_stray_grads()
            self.arena.allreduce_sum(self.group)
            self.opt.step(grads_are_summed=True)
        self.optimizer_steps += 1

    def flush(self) -> Optional[int]:
        """Step on a partial window (end of an epoch whose length is not a multiple of the window, as Lightning does)."""
        if self.micro == 0:
            return None
        done = self.micro
        self.micro = 0
        self._exchange_and_step(exchanged=False)     # the window's micro-batches all ran under no_sync()
        return done
