"""Whole-step CUDA graph for the training loop (launch-bound inner loop → one graph launch).

A MedNeXt-S step is ~1 400 kernel launches (ours + loss/optimizer/repack elementwise ops); at 45 ms/step
the host launch path is already ~10 % of the step.  ``GraphedTrainStep`` captures
forward + loss + backward + gradient all-reduce + optimizer step once (all pcb200 kernels are enqueued on
the current stream, so they are capturable as-is) and replays it per step with static input buffers.
Semantics are those of the eager step in ``bench.py`` / ``training/lightning/model.py:863-910``.
"""

from __future__ import annotations

from typing import Callable

import torch

from .ddp import FlatGradArena


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, loss_fn: Callable, optimizer: torch.optim.Optimizer,
                 arena: FlatGradArena, example_input: torch.Tensor, example_target: torch.Tensor, warmup: int = 3):
        self.model, self.loss_fn, self.opt, self.arena = model, loss_fn, optimizer, arena
        self.x = example_input.clone()
        self.t = example_target.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._step()

    def _step(self) -> torch.Tensor:
        self.arena.zero()
        loss = self.loss_fn(self.model(self.x), self.t)
        loss.backward()
        self.arena.allreduce()
        self.opt.step()
        return loss

    def __call__(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        self.x.copy_(x, non_blocking=True)
        self.t.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.loss
