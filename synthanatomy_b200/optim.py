"""Adam with the library's fused kernel; drop-in for ``torch.optim.Adam(params, lr)`` as the reference uses it
(/root/reference/run_vqvae.py:82, run_transformer.py:109: no weight decay, no amsgrad)."""
from __future__ import annotations

import torch

from . import ops


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                if torch.is_tensor(st["step"]):     # state restored from a torch.optim.Adam checkpoint (tensor step)
                    st["step"] = int(st["step"].item())
                st["step"] += 1
                ops.adam_step(p.data, p.grad.contiguous(), st["exp_avg"], st["exp_avg_sq"], group["lr"], b1, b2,
                              group["eps"], st["step"])
        return loss
