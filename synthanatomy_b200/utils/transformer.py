"""Mirror of the batch preparation the Performer's callers apply, /root/reference/src/utils/transformer.py:239-317
(integer reshaping / gathering only -- no arithmetic): flatten the token grid, reorder by the ordering's index
sequence, left-pad BOS (= vocab_size), split into input / target."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _to(t, device, non_blocking):
    return t.to(device=device, non_blocking=non_blocking) if device is not None else t


def prepare_batch(batch, index_sequence, vocab_size, conditionings=None, device=None, non_blocking=False):
    encoded = batch["quantization"]
    encoded = encoded.reshape(encoded.shape[0], -1)
    encoded = encoded[:, index_sequence]
    encoded = F.pad(encoded, (1, 0), "constant", vocab_size)
    encoded = encoded.long()
    conditioned = None
    if conditionings:
        conditioned = []
        for label in conditionings:
            c = batch[label]
            if len(c.shape) == 1:
                c = c[..., None]
            conditioned.append(_to(c.long(), device, non_blocking))
    x_input = _to(encoded[:, :-1], device, non_blocking)
    x_target = _to(encoded[:, 1:], device, non_blocking)
    return (x_input, conditioned), x_target


def prepare_inference_batch(batch, num_embeddings, conditionings=None, device=None, non_blocking=False):
    no_samples = batch["quantization"].shape[0]
    initial = torch.from_numpy(np.repeat(np.array([[num_embeddings]]), no_samples, axis=0)).long()
    conditioned = None
    if conditionings:
        conditioned = []
        for label in conditionings:
            c = batch[label]
            if len(c.shape) == 1:
                c = c[..., None]
            conditioned.append(_to(c.long(), device, non_blocking))
    return (_to(initial, device, non_blocking), conditioned), _to(initial, device, non_blocking)
