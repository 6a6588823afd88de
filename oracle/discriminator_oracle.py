"""CPU oracle for the PatchGAN discriminator (TEST INFRASTRUCTURE -- never imported by the product).

A functional restatement of ``/root/reference/src/networks/discriminator/baseline.py:21-88`` that works on a plain
``state_dict`` (same keys as the reference ``BaselineDiscriminator``: ``main.<i>.weight`` ...), so that it runs on the
GPU box where ``/root/reference`` does not exist.

Parity status: PINNED.  ``oracle/make_golden_discriminator.py`` runs the unmodified reference class in the build
container and stores its state, outputs, gradients and updated BatchNorm buffers in ``tests/golden/discriminator.npz``;
``tests/test_discriminator_oracle.py`` checks this restatement against that file.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

SLOPE = 0.2        # nn.LeakyReLU(0.2, True), baseline.py:44,62,78
BN_EPS = 1e-5      # nn.BatchNorm3d defaults
BN_MOMENTUM = 0.1


def layer_plan(state: Dict[str, torch.Tensor]) -> List[Tuple[int, int, bool, bool]]:
    """[(index of the conv inside ``main``, stride, followed by BatchNorm, followed by LeakyReLU)] in execution order,
    recovered from the key layout of baseline.py:41-81: conv(s2)+act, (n_layers-1) x conv(s2)+bn+act, conv(s1)+bn+act,
    conv(s1)."""
    conv_idx = sorted({int(k.split(".")[1]) for k in state if k.endswith(".weight") and state[k].dim() == 5})
    plan = []
    for j, i in enumerate(conv_idx):
        has_bn = f"main.{i + 1}.running_mean" in state
        last = j == len(conv_idx) - 1
        stride = 1 if j >= len(conv_idx) - 2 else 2
        plan.append((i, stride, has_bn, not last))
    return plan


def forward(state: Dict[str, torch.Tensor], x: torch.Tensor, training: bool = True) -> torch.Tensor:
    """``BaselineDiscriminator.forward`` (baseline.py:86-88) on ``state``; BatchNorm buffers in ``state`` are updated in
    place when ``training`` (running statistics with momentum 0.1 and ``num_batches_tracked``)."""
    h = x
    for i, stride, has_bn, has_act in layer_plan(state):
        h = F.conv3d(h, state[f"main.{i}.weight"], state.get(f"main.{i}.bias"), stride=stride, padding=1)
        if has_bn:
            b = i + 1
            if training:
                state[f"main.{b}.num_batches_tracked"] += 1
            h = F.batch_norm(h, state[f"main.{b}.running_mean"], state[f"main.{b}.running_var"], state[f"main.{b}.weight"],
                             state[f"main.{b}.bias"], training, BN_MOMENTUM, BN_EPS)
        if has_act:
            h = F.leaky_relu(h, SLOPE)
    return h
