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
    """``split_collective=False``: the whole step (gradient all-reduce included) is ONE graph.
    ``split_collective=True`` (data-parallel default): two graphs around an eagerly launched all-reduce —
    graph A = zero grads + forward + loss + backward, then ``dist.all_reduce`` over the flat arena on the same
    stream, then graph B = 1/world scale + optimizer step — three host launches per step instead of ~1 400, and
    no NCCL call inside a capture."""

    def __init__(self, model: torch.nn.Module, loss_fn: Callable, optimizer: torch.optim.Optimizer,
                 arena: FlatGradArena, example_input: torch.Tensor, example_target: torch.Tensor, warmup: int = 3,
                 split_collective: bool = False, group=None):
        self.model, self.loss_fn, self.opt, self.arena = model, loss_fn, optimizer, arena
        self.split, self.group = bool(split_collective), group
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
        self.graph_opt = None
        if not self.split:
            with torch.cuda.graph(self.graph):
                self.loss = self._step()
            return
        with torch.cuda.graph(self.graph):
            self.loss = self._fwd_bwd()
        self.arena.allreduce_sum(self.group)
        self.graph_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_opt, pool=self.graph.pool()):
            self.arena.scale_mean(self.group)
            self.opt.step()

    def _fwd_bwd(self) -> torch.Tensor:
        self.arena.zero()
        loss = self.loss_fn(self.model(self.x), self.t)
        loss.backward()
        return loss

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
        if self.graph_opt is not None:
            self.arena.allreduce_sum(self.group)
            self.graph_opt.replay()
        return self.loss
