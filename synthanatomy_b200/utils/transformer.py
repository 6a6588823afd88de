"""Host-side batch preparation for the Performer: the tensors the reference's training / inference engines hand to the
network (/root/reference/src/utils/transformer.py:239-317), restated.  Integer reshaping only -- no arithmetic:

    tokens  = grid.reshape(B, -1)[:, ordering]              # raster -> sequence order
    x_input = [BOS, tokens[:-1]],  x_target = tokens        # BOS = vocab_size, i.e. one past the largest code

``synthanatomy_b200.utils.tokens.prepare_batch_device`` produces the same pair with one gather kernel from the uint16
grid; this module is the host form (CPU tensors in, optional transfer at the end) that the tests pin to the reference.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch


def _place(t: torch.Tensor, device, non_blocking: bool) -> torch.Tensor:
    return t if device is None else t.to(device=device, non_blocking=non_blocking)


def _conditionings(batch: Dict, labels: Optional[Sequence[str]], device, non_blocking: bool) -> Optional[List[torch.Tensor]]:
    """every conditioning as an int64 column [B, 1] (a 1-D label vector gains the trailing axis); None when there are none"""
    if not labels:
        return None
    cols = []
    for name in labels:
        c = batch[name]
        c = c.unsqueeze(-1) if c.dim() == 1 else c
        cols.append(_place(c.long(), device, non_blocking))
    return cols


def prepare_batch(batch: Dict, index_sequence, vocab_size: int, conditionings: Optional[Sequence[str]] = None, device=None,
                  non_blocking: bool = False) -> Tuple[Tuple[torch.Tensor, Optional[List[torch.Tensor]]], torch.Tensor]:
    """((x_input, conditioned), x_target) for one training / evaluation batch"""
    grid = batch["quantization"]
    tokens = grid.reshape(grid.shape[0], -1)[:, index_sequence].long()
    bos = torch.full((tokens.shape[0], 1), vocab_size, dtype=torch.long, device=tokens.device)
    x_input = torch.cat((bos, tokens[:, :-1]), dim=1)
    return (_place(x_input, device, non_blocking), _conditionings(batch, conditionings, device, non_blocking)), \
        _place(tokens, device, non_blocking)


def prepare_inference_batch(batch: Dict, num_embeddings: int, conditionings: Optional[Sequence[str]] = None, device=None,
                            non_blocking: bool = False):
    """sampling starts from the BOS token alone: ((prefix [B, 1], conditioned), prefix)"""
    prefix = torch.full((batch["quantization"].shape[0], 1), num_embeddings, dtype=torch.long)
    prefix = _place(prefix, device, non_blocking)
    return (prefix, _conditionings(batch, conditionings, device, non_blocking)), prefix
