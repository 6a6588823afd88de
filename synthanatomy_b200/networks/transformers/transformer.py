"""Mirror of /root/reference/src/networks/transformers/transformer.py:9-104 (``TransformerBase``): the autoregressive
sampling loop the inference mode drives (src/inferer/transformer.py:63-71).  Same methods, arguments and results;
the per-token forward runs through the B200 kernels of the subclass."""
from __future__ import annotations

from typing import Any, Optional

import numpy as np
import torch


class TransformerBase(torch.nn.Module):
    """Abstract class for transformers."""

    @staticmethod
    def _top_k_logits(logits: torch.Tensor, k: int) -> torch.Tensor:
        v, _ = torch.topk(logits, k)
        out = logits.clone()
        out[out < v[:, [-1]]] = -float("Inf")
        return out

    @torch.no_grad()
    def sample_next_index(self, x: torch.Tensor, conditioning: torch.Tensor = None, temperature: float = 1.0,
                          sample: bool = True, top_k: Optional[int] = None) -> torch.Tensor:
        """transformer.py:20-56: full forward over the prefix, softmax of the last position, multinomial / top-1."""
        self.eval()
        logits = self(x, conditioning)
        logits = logits[:, -1, :] / temperature
        if top_k is not None:
            logits = self._top_k_logits(logits, top_k)
        probs = torch.softmax(logits, dim=-1)
        if sample:
            ix = torch.multinomial(probs, num_samples=1)
        else:
            _, ix = torch.topk(probs, k=1, dim=-1)
        return ix

    @torch.no_grad()
    def sample(self, prefix: torch.Tensor, conditioning: torch.Tensor = None, temperature: float = 1.0,
               sample: bool = True, top_k: Optional[int] = None) -> torch.Tensor:
        """transformer.py:58-101: prod(ordering.dimensions) steps, then undo the ordering and reshape to the grid."""
        steps = int(np.prod(self.ordering.dimensions))
        x = prefix
        for _ in range(steps):
            ix = self.sample_next_index(x, conditioning=conditioning, temperature=temperature, sample=sample, top_k=top_k)
            x = torch.cat((x, ix), dim=1)
        x = x[:, prefix.shape[1]:]
        x = x[:, self.ordering.get_revert_sequence_ordering()]
        x = x.reshape(x.shape[0], *self.ordering.dimensions)
        x = torch.squeeze(x, 1)
        return x

    def forward(self, x: torch.Tensor) -> Any:  # pragma: no cover - abstract
        raise NotImplementedError
