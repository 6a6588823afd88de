"""CPU oracle for the spectral reconstruction loss of the README run (TEST INFRASTRUCTURE -- never imported by the
product; the CUDA path it checks is synthanatomy_b200.losses.JukeboxLoss,
tests/test_gpu_losses.py).

``jukebox_loss`` restates ``JukeboxLoss.forward`` (/root/reference/src/losses/vqvae/vqvae.py:522-640): orthonormal FFT
over the channel and spatial axes, amplitude ``sqrt(re^2 + im^2)``, mean squared amplitude difference times
``fft_factor``, plus the pixel MSE and the quantisation losses.

Parity status: PINNED.  ``oracle/make_golden_losses.py`` runs the unmodified reference class (stubs only for the
``lpips`` import and the TensorBoard enum, neither of which this loss executes) and stores inputs, value and gradient in
``tests/golden/jukebox.npz``; ``tests/test_losses_oracle.py`` checks this restatement against it.
"""
from __future__ import annotations

from typing import Sequence

import torch


def fft_amplitude(images: torch.Tensor, dimensions: int = 3) -> torch.Tensor:
    """vqvae.py:617-630 with the default fft_kwargs (:585-589): axes 1 .. dimensions + 1, norm="ortho" """
    spec = torch.fft.fftn(images, dim=tuple(range(1, dimensions + 2)), norm="ortho")
    return torch.sqrt(spec.real ** 2 + spec.imag ** 2)


def jukebox_loss(reconstruction: torch.Tensor, y: torch.Tensor, quantization_losses: Sequence[torch.Tensor] = (),
                 dimensions: int = 3, fft_factor: float = 1.0, include_pixel_loss: bool = True) -> torch.Tensor:
    y = y.float()
    pred = reconstruction.float()
    loss = torch.mean((fft_amplitude(pred, dimensions) - fft_amplitude(y, dimensions)) ** 2) * fft_factor      # :598-599
    if include_pixel_loss:
        loss = loss + torch.mean((pred - y) ** 2)                                                              # :603-607
    for q in quantization_losses:                                                                              # :609-616
        loss = loss + q.float()
    return loss
