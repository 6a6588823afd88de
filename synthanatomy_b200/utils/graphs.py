"""CUDA-graph replay of a training step (forward + loss + backward) of the drop-in networks.

At the README's own Performer size (latent grid 10 x 14 x 10 = 1400 tokens) a step is ~1000 kernel launches of a few
microseconds each: the GPU waits for the host.  Capturing the step once and replaying it removes the per-launch host cost;
the kernels, their order and their arithmetic are exactly those of the eager step.

    step = GraphedTrainStep(model, loss_fn, optimizer, (x_example,), y_example, before_step=model.check_redraw_projections)
    for x, y in loader:
        loss = step(x, target=y)          # copies into the static inputs, replays, then runs optimizer.step() eagerly

What stays OUTSIDE the graph, on purpose: the optimiser step (its bias correction is computed on the host from the step
count -- four multi-tensor launches), and anything passed as `before_step` (the Performer's projection redraw: host QR on
a prefetch thread + H2D copies into the static projection buffers the captured kernels read).

Single-GPU (or one replica per process without DistributedDataParallel): the reducer's bucket all-reduce is not captured.
PyTorch plumbing only (torch.cuda.CUDAGraph, streams); no arithmetic.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, loss_fn: Callable, optimizer: torch.optim.Optimizer,
                 example_inputs: Sequence[torch.Tensor], example_target: torch.Tensor, warmup: int = 3,
                 before_step: Optional[Callable[[], None]] = None, forward: Optional[Callable] = None):
        """forward(model, *inputs) -> what loss_fn(output, target) takes (default: model(*inputs))"""
        if not example_target.is_cuda:
            raise RuntimeError("synthanatomy_b200: CUDA tensors only -- there is no CPU fallback")
        self.model, self.loss_fn, self.optimizer, self.before_step = model, loss_fn, optimizer, before_step
        self.forward = forward or (lambda m, *a: m(*a))
        self.inputs = [t.clone() for t in example_inputs]
        self.target = example_target.clone()
        # the projection redraw (host logic, H2D copies) must not run inside the captured region
        stack = getattr(model, "performer", None)
        self._stack = stack if hasattr(stack, "auto_check_redraw") else None
        auto = self._stack.auto_check_redraw if self._stack is not None else None
        if self._stack is not None:
            self._stack.auto_check_redraw = False
        try:
            # graphs of earlier eager steps (kept alive by a stored loss, say) would keep the parameters' gradient
            # accumulators on the stream those steps ran on, which invalidates the capture below
            import gc
            gc.collect()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):                     # allocator, lazy kernel attributes, workspaces
                    if before_step is not None:
                        before_step()
                    optimizer.zero_grad(set_to_none=True)
                    self.loss_fn(self.forward(model, *self.inputs), self.target).backward()
                    optimizer.step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            optimizer.zero_grad(set_to_none=True)           # the captured backward creates the (static) .grad tensors
            with torch.cuda.graph(self.graph):
                self.loss = self.loss_fn(self.forward(model, *self.inputs), self.target)
                self.loss.backward()
        finally:
            if self._stack is not None:
                self._auto = auto
        # from here on the module keeps auto_check_redraw off: the redraw runs in __call__, before the replay

    def __call__(self, *inputs: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        for dst, src in zip(self.inputs, inputs):
            dst.copy_(src, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        if self.before_step is not None:
            self.before_step()
        self.graph.replay()
        self.optimizer.step()
        return self.loss

    def release(self) -> None:
        """drop the graph (and its private memory pool); restores the module's own redraw trigger"""
        self.graph = None
        if self._stack is not None:
            self._stack.auto_check_redraw = self._auto
