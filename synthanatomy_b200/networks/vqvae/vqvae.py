"""Plugin contract of the VQ-VAE family: mirror of the reference's abstract ``VQVAEBase``
(/root/reference/src/networks/vqvae/vqvae.py:8-140): same method names, argument meaning and return types, so
callers (inferers, handlers, trainers) cannot tell the implementations apart."""
from __future__ import annotations

import abc
from typing import Dict, List, Sequence, Tuple, Union

import torch
import torch.nn as nn


class VQVAEBase(nn.Module, metaclass=abc.ABCMeta):
    @abc.abstractmethod
    def forward(self, images: torch.Tensor) -> Dict[str, List[torch.Tensor]]:
        """-> {"reconstruction": [Tensor], "quantization_losses": [Tensor]}"""

    @abc.abstractmethod
    def encode(self, images: torch.Tensor) -> List[torch.Tensor]:
        """images -> list of encodings to be passed to ``quantize``"""

    @abc.abstractmethod
    def quantize(self, encodings: List[torch.Tensor]) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
        """encodings -> (quantizations, quantization losses)"""

    @abc.abstractmethod
    def decode(self, quantizations: List[torch.Tensor]) -> torch.Tensor:
        """quantizations -> reconstruction"""

    @abc.abstractmethod
    def index_quantize(self, images: torch.Tensor) -> List[torch.Tensor]:
        """images -> list of LongTensor code indices (input of the autoregressive prior)"""

    @abc.abstractmethod
    def decode_samples(self, embedding_indices: List[torch.Tensor]) -> torch.Tensor:
        """code indices -> images"""

    @abc.abstractmethod
    def get_ema_decay(self) -> Sequence[float]:
        ...

    @abc.abstractmethod
    def set_ema_decay(self, decay: Union[Sequence[float], float]) -> Sequence[float]:
        ...

    @abc.abstractmethod
    def get_commitment_cost(self) -> Sequence[float]:
        ...

    @abc.abstractmethod
    def set_commitment_cost(self, commitment_factor: Union[Sequence[float], float]) -> Sequence[float]:
        ...

    @abc.abstractmethod
    def get_perplexity(self) -> Sequence[float]:
        ...

    @abc.abstractmethod
    def get_last_layer(self) -> nn.parameter.Parameter:
        ...
