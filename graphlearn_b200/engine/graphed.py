"""Whole-step CUDA-graph capture for arbitrary (static-shape) models.

``FastSageTrainer`` hand-schedules GraphSAGE; every other model of the zoo (EgoGAT towers, GIN, RGCN, ...) trains through
autograd over the fused kernels, where a step is hundreds of small launches and the host - not the GPU - is the bottleneck.
:class:`GraphedTrainStep` removes the host from the loop the way the reference's session.run removes Python from a TF1
step: forward + loss + backward + optimiser step are captured ONCE into a CUDA graph over static input buffers; a step is
then "copy the batch's id tensors into the buffers, replay".  Batches whose shapes differ from the captured ones (the
short tail batch of an epoch) run eagerly.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch


class GraphedTrainStep(object):
    def __init__(self, loss_fn: Callable[[Dict[str, torch.Tensor]], torch.Tensor], optimizer: torch.optim.Optimizer,
                 example: Dict[str, torch.Tensor], warmup: int = 3, grad_hook: Optional[Callable[[], None]] = None):
        """loss_fn(inputs) -> scalar loss (must be free of host synchronisation and of data-dependent shapes);
        ``optimizer`` must be capturable (e.g. ``torch.optim.Adam(..., capturable=True)``); ``example``: one batch of
        inputs - its shapes become the static shapes; ``grad_hook`` runs between backward and the optimiser step
        (gradient all-reduce)."""
        self.loss_fn, self.opt, self.grad_hook = loss_fn, optimizer, grad_hook
        self.static = {k: v.clone() for k, v in example.items()}
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.loss: Optional[torch.Tensor] = None
        self.replays = self.eager_steps = 0
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(max(1, warmup)):
                self._eager(self.static)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(g):
            loss = self.loss_fn(self.static)
            loss.backward()
            if self.grad_hook is not None:
                self.grad_hook()
            self.opt.step()
            self.loss = loss.detach()
        self.graph = g

    def _eager(self, inputs) -> torch.Tensor:
        self.opt.zero_grad(set_to_none=True)
        loss = self.loss_fn(inputs)
        loss.backward()
        if self.grad_hook is not None:
            self.grad_hook()
        self.opt.step()
        return loss.detach()

    def __call__(self, inputs: Dict[str, torch.Tensor]) -> torch.Tensor:
        """one training step; returns the (device) loss tensor of this step"""
        if self.graph is None or any(tuple(inputs[k].shape) != tuple(v.shape) for k, v in self.static.items()):
            self.eager_steps += 1
            return self._eager(inputs)
        for k, v in self.static.items():
            v.copy_(inputs[k], non_blocking=True)
        self.graph.replay()
        self.replays += 1
        return self.loss
