"""One adversarial training iteration without Ignite: the arithmetic of ``AdversarialTrainer._iteration``
(/root/reference/src/engines/trainer.py:122-262) as a plain function over the two networks, their optimisers and losses.

    generator:      recon = L_rec(G(x), x);  g = L_g(D(G(x).reconstruction));  w = adaptive weight;  (recon + w g).backward(); step
    discriminator:  d = w * L_d(D(G(x).detach()), D(x));  d.backward(); step

Host logic only (device-agnostic): the networks are whatever modules are passed in -- on a B200 the drop-in VQ-VAE and
discriminator of this package, whose kernels do all the arithmetic.  ``torch.autocast(bf16)`` replaces the reference's
fp16 autocast + GradScaler pair (bf16 needs no loss scaling).
"""
from __future__ import annotations

import contextlib
from typing import Callable, Dict, Optional

import torch


def adaptive_adversarial_weight(g_network, reconstruction_loss: torch.Tensor, generator_loss: torch.Tensor, global_step: int,
                                enabled: bool, threshold: int = 0, value: float = 1.0):
    """trainer.py:264-289: ratio of the gradient norms of the two loss terms at the generator's last layer, clamped to
    [0, 1e4]; ``value`` during the first ``threshold`` epochs; 1 when disabled."""
    if not enabled:
        return 1
    last = g_network.get_last_layer()
    nll_grads = torch.autograd.grad(reconstruction_loss, last, retain_graph=True)[0]
    g_grads = torch.autograd.grad(generator_loss, last, retain_graph=True)[0]
    weight = torch.clamp(torch.norm(nll_grads) / (torch.norm(g_grads) + 1e-4), 0.0, 1e4).detach()
    if global_step < threshold:
        weight = value
    return weight


def adversarial_iteration(inputs: torch.Tensor, targets: torch.Tensor, g_network, d_network, g_optimizer, d_optimizer,
                          recon_loss_function: Callable, g_loss_function: Callable, d_loss_function: Callable, epoch: int = 0,
                          use_adversarial_adaptive_weight: bool = False, adaptive_adversarial_weight_threshold: int = 0,
                          adaptive_adversarial_weight_value: float = 1.0, amp: bool = False,
                          amp_dtype: Optional[torch.dtype] = torch.bfloat16) -> Dict:
    ctx = (lambda: torch.autocast(inputs.device.type, dtype=amp_dtype)) if amp else contextlib.nullcontext
    retain = contextlib.nullcontext
    if use_adversarial_adaptive_weight:
        # the adaptive weight differentiates the generator's graph three times; the drop-in VQ-VAE must keep its activations
        from .networks.vqvae import b200 as _b200
        retain = _b200.retain_activations

    # ---- generator (trainer.py:157-213)
    g_network.train()
    g_optimizer.zero_grad(set_to_none=True)
    with retain():
        with ctx():
            g_predictions = g_network(inputs)
            logits_fake = d_network(g_predictions["reconstruction"][0].float().contiguous())
            reconstruction_loss = recon_loss_function(g_predictions, targets).mean()
            generator_loss = g_loss_function(logits_fake).mean()
            adversarial_weight = adaptive_adversarial_weight(g_network, reconstruction_loss, generator_loss, epoch,
                                                             use_adversarial_adaptive_weight,
                                                             adaptive_adversarial_weight_threshold,
                                                             adaptive_adversarial_weight_value)
            generator_loss = reconstruction_loss + generator_loss * adversarial_weight
    # outside `retain`: this last differentiation releases the stacks' saved activations (tens of GB at README size)
    # as the reference's backward() does, instead of keeping them alive through the discriminator step
    generator_loss.backward()
    g_optimizer.step()

    # ---- discriminator (trainer.py:215-251): gradients the generator pass left on it are dropped first
    d_network.train()
    d_network.zero_grad(set_to_none=True)
    with ctx():
        logits_fake = d_network(g_predictions["reconstruction"][0].float().contiguous().detach())
        logits_real = d_network(inputs.contiguous().detach())
        d_loss = d_loss_function(logits_fake, logits_real).mean() * adversarial_weight
    d_loss.backward()
    d_optimizer.step()

    return {"image": inputs, "label": targets, "pred": g_predictions, "loss": reconstruction_loss.item(), "reals": inputs,
            "fakes": g_predictions, "g_loss": generator_loss.item(), "d_loss": d_loss.item()}
