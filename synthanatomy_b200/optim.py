"""Adam with the library's fused kernel; drop-in for ``torch.optim.Adam(params, lr)`` as the reference uses it
(/root/reference/run_vqvae.py:82, run_transformer.py:109: no weight decay, no amsgrad).

One optimiser step is a handful of multi-tensor launches (``sa_adam_multi``: 64 parameter tensors per launch) instead of
one launch per parameter.  The ``state_dict`` is interchangeable with ``torch.optim.Adam``'s in both directions
(same param_group keys, ``step`` / ``exp_avg`` / ``exp_avg_sq`` state; tests/test_checkpoints.py)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, ops


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        # the param_groups carry every key torch.optim.Adam's own groups have, so that a state_dict saved here can be
        # loaded into torch.optim.Adam by the reference's entry points (and stepped) and the other way round
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False)
        super().__init__(params, defaults)
        self._check_groups()

    def _check_groups(self):
        for group in self.param_groups:
            if group.get("weight_decay", 0) != 0 or group.get("amsgrad", False) or group.get("maximize", False):
                raise NotImplementedError("synthanatomy_b200.optim.Adam implements the reference's configuration only "
                                          "(run_vqvae.py:82, run_transformer.py:109: weight_decay=0, amsgrad=False, "
                                          "maximize=False); no fallback")

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._check_groups()

    @staticmethod
    def _step_of(st) -> int:
        s = st["step"]
        return int(s.item()) if torch.is_tensor(s) else int(s)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._check_groups()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            by_step = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("synthanatomy_b200: CUDA tensors only -- there is no CPU fallback")
                st = self.state[p]
                if not st:
                    st["step"] = torch.tensor(0.0)          # a CPU tensor, as torch.optim.Adam keeps it
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                k = self._step_of(st) + 1
                st["step"] = torch.tensor(float(k))
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                by_step.setdefault(k, []).append((p, g, st["exp_avg"], st["exp_avg_sq"]))
            for k, items in by_step.items():
                n = len(items)
                arrs = [(C.c_void_p * n)() for _ in range(4)]
                sizes = (C.c_int64 * n)()
                for i, tup in enumerate(items):
                    assert tup[0].dtype == torch.float32 and tup[0].is_contiguous(), "fp32 contiguous parameters only"
                    for a, t in zip(arrs, tup):
                        a[i] = t.data_ptr()
                    sizes[i] = tup[0].numel()
                _lib.check(ops.lib().sa_adam_multi(n, arrs[0], arrs[1], arrs[2], arrs[3], sizes, float(group["lr"]),
                                                   float(b1), float(b2), float(group["eps"]), int(k), ops._stream()),
                           "sa_adam_multi")
                del items
        return loss
