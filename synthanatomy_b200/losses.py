"""Losses on the hot path: mirrors of the reference's ``MSELoss`` for the VQ-VAE
(/root/reference/src/losses/vqvae/vqvae.py:14-71): ``mse(reconstruction, y) + sum(quantization_losses)``, and of its
``CELoss`` for the Performer (/root/reference/src/losses/transformer/transformer.py:10-36)."""
from __future__ import annotations

from typing import Dict, List

import torch
from torch.nn.modules.loss import _Loss

from . import ops, pf_ops


class _MSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        p = pred.detach().float().contiguous()
        t = target.detach().float().contiguous()
        sse = torch.zeros((), device=p.device, dtype=torch.float32)
        ops.mse_fwd_bwd(p, t, sse, None, 1.0)
        ctx.save_for_backward(p, t)
        return sse / p.numel()

    @staticmethod
    def backward(ctx, g):
        p, t = ctx.saved_tensors
        grad = torch.empty_like(p)
        ops.mse_fwd_bwd(p, t, None, grad, 2.0 / p.numel(), g.float().contiguous())
        return grad, None


def mse_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """F.mse_loss(pred, target) (mean reduction) through the library's fused kernel."""
    return _MSEFn.apply(pred, target)


class MSELoss(_Loss):
    def __init__(self, size_average: bool = None, reduce: bool = None, reduction: str = "mean"):
        super().__init__(size_average, reduce, reduction)
        if reduction not in ("sum", "mean"):
            raise ValueError("Reduction must be either 'sum' or 'mean'")
        self.summaries: Dict = {"scalar": dict()}

    def forward(self, network_output: Dict[str, List[torch.Tensor]], y: torch.Tensor) -> torch.Tensor:
        y_pred = network_output["reconstruction"][0]
        q_losses = network_output["quantization_losses"]
        loss = mse_loss(y_pred, y)     # the reference ignores `reduction` here as well (F.mse_loss default)
        self.summaries["scalar"]["Loss-MSE-Reconstruction"] = loss
        for idx, q_loss in enumerate(q_losses):
            q_loss = q_loss.float()
            self.summaries["scalar"][f"Loss-MSE-VQ{idx}_Commitment_Cost"] = q_loss
            loss = loss + q_loss
        return loss

    def get_summaries(self) -> Dict[str, torch.Tensor]:
        return self.summaries


class _CEFn(torch.autograd.Function):
    """F.cross_entropy(input [B, V, N], target [B, N]) with one fused kernel per direction (no log-softmax tensor)."""

    @staticmethod
    def forward(ctx, y_pred, y, reduction):
        if not y_pred.is_cuda:
            raise RuntimeError("synthanatomy_b200: CUDA tensors only -- there is no CPU fallback")
        logits = y_pred.detach().float().transpose(1, 2).contiguous()      # [B, N, V]; a view for TransformerTrainingInferer output
        B, N, V = logits.shape
        target = y.detach().long().contiguous().view(-1)
        loss_sum = torch.zeros((1,), device=logits.device, dtype=torch.float32)
        pf_ops.ce_fwd_bwd(logits.view(B * N, V), target, 0.0, None, loss_sum, None)
        ctx.save_for_backward(logits, target)
        ctx.scale = 1.0 / (B * N) if reduction == "mean" else 1.0
        return loss_sum[0] * ctx.scale

    @staticmethod
    def backward(ctx, g):
        logits, target = ctx.saved_tensors
        B, N, V = logits.shape
        dl = torch.empty_like(logits)
        pf_ops.ce_fwd_bwd(logits.view(B * N, V), target, ctx.scale, g.float().contiguous().view(1), None, dl.view(B * N, V))
        return dl.transpose(1, 2), None, None


class CELoss(_Loss):
    def __init__(self, weight=None, size_average: bool = None, reduce: bool = None, reduction: str = "mean"):
        super().__init__(size_average, reduce, reduction)
        if reduction not in ("sum", "mean"):
            raise ValueError("Reduction must be either 'sum' or 'mean'")
        if weight is not None:
            raise NotImplementedError("class weights are not implemented (the reference passes none); no fallback")
        self.summaries: Dict = {"scalar": dict()}

    def forward(self, y_pred: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        loss = _CEFn.apply(y_pred, y, self.reduction)
        self.summaries["scalar"]["Loss-CE-Prediction"] = loss
        return loss

    def get_summaries(self) -> Dict[str, torch.Tensor]:
        return self.summaries


# ------------------------------------------------------------------------------------------------
# adversarial criteria on the discriminator's patch map (/root/reference/src/losses/adversarial/adversarial.py:11-122).
# The map is ~7e4 values at the README size: a handful of elementwise torch ops, device-agnostic host logic.
# ------------------------------------------------------------------------------------------------
ADVERSARIAL_CRITERIA = ("vanilla", "hinge", "least_square")          # src/losses/adversarial/utils.py


def adversarial_criterion(name: str):
    """per-element loss(logits, is_real); note the reference's naming: "vanilla" is the relu hinge, "hinge" the softplus
    form (adversarial.py:85-120)"""
    sign = lambda is_real: -1.0 if is_real else 1.0                   # noqa: E731
    if name == "vanilla":
        return lambda logits, is_real: torch.relu(1.0 + sign(is_real) * logits)
    if name == "hinge":
        return lambda logits, is_real: torch.nn.functional.softplus(sign(is_real) * logits)
    if name == "least_square":
        return lambda logits, is_real: (logits - (1.0 if is_real else 0.0)) ** 2
    raise ValueError(f"Unknown adversarial loss. Available losses are {list(ADVERSARIAL_CRITERIA)} but received {name}")


class AdversarialLoss(_Loss):
    """generator side (is_discriminator=False): weight * mean(criterion(D(fake), real=True));
    discriminator side: weight * 0.5 * (mean(criterion(D(fake), False)) + mean(criterion(D(real), True)))"""

    def __init__(self, criterion: str = "least_square", is_discriminator: bool = True, weight=None, size_average: bool = None,
                 reduce: bool = None, reduction: str = "mean"):
        super().__init__(size_average, reduce, reduction)
        if reduction not in ("sum", "mean"):
            raise ValueError("Reduction must be either 'sum' or 'mean'")
        self.criterion = criterion
        self.is_discriminator = is_discriminator
        self.criterion_function = adversarial_criterion(criterion)
        self._weight = weight
        self.summaries: Dict = {"scalar": dict()}

    def forward(self, logits_fake: torch.Tensor, logits_real: torch.Tensor = None) -> torch.Tensor:
        side = "Discriminator" if self.is_discriminator else "Generator"
        loss = torch.mean(self.criterion_function(logits_fake.float(), not self.is_discriminator))
        self.summaries["scalar"][f"Loss-Adversarial_{side}-Reconstruction"] = loss
        if self.is_discriminator:
            loss_real = torch.mean(self.criterion_function(logits_real.float(), True))
            self.summaries["scalar"]["Loss-Adversarial_Discriminator-Originals"] = loss_real
            loss = 0.5 * (loss + loss_real)
        return self._weight * loss

    def get_summaries(self) -> Dict[str, torch.Tensor]:
        return self.summaries

    def get_weight(self) -> float:
        return self._weight

    def set_weight(self, weight: float) -> float:
        self._weight = weight
        return self.get_weight()


def get_discriminator_loss(config: dict) -> AdversarialLoss:
    """src/losses/adversarial/configure.py:10-22"""
    adversarial_criterion(config["discriminator_loss"])
    return AdversarialLoss(criterion=config["discriminator_loss"], is_discriminator=True, weight=0.005)


def get_generator_loss(config: dict) -> AdversarialLoss:
    """src/losses/adversarial/configure.py:25-37"""
    adversarial_criterion(config["generator_loss"])
    return AdversarialLoss(criterion=config["generator_loss"], is_discriminator=False, weight=0.005)
