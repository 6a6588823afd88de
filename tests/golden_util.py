"""Helpers to load tests/golden/*.npz (made by oracle/make_golden.py from the unmodified reference)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def vqvae_case(name):
    """-> (cfg kwargs, sd0 (torch), blob)"""
    blob = load(name)
    n_levels, ch, n_embed, embed_dim, batch, *vol = [int(v) for v in blob["cfg"]]
    cfg = dict(n_levels=n_levels, downsample_parameters=((4, 2, 1, 1),) * n_levels,
               upsample_parameters=((4, 2, 1, 0, 1),) * n_levels, n_embed=n_embed, embed_dim=embed_dim,
               n_channels=ch, n_res_channels=ch, n_res_layers=3, vq_decay=0.5, commitment_cost=0.25)
    sd = {k[4:]: torch.from_numpy(v) for k, v in blob.items() if k.startswith("sd0/")}
    sd["quantizer.0.impl.embedding.weight"] = sd["quantizer.0.impl.weight"]
    sd["quantizer.0.impl.embed_avg"] = sd["quantizer.0.impl.weight"].clone()
    return cfg, sd, blob


def vq_inputs(blob, tag):
    """Rebuild the latent / codebook pair of a vq_cfg3 case exactly as oracle/make_golden.py:vq_cases did."""
    z = torch.from_numpy(blob["z"])
    W = torch.from_numpy(blob["W"])
    D = W.shape[1]
    if tag == "plain":
        return z, W
    if tag == "dup":
        W2 = W.clone(); W2[7] = W2[3]; W2[100] = W2[3]; W2[2047] = W2[0]
        return z, W2
    if tag == "exact":
        sel = torch.from_numpy(blob["exact/sel"].astype(np.int64))
        return W[sel].reshape(2, 10, 14, 10, D).permute(0, 4, 1, 2, 3).contiguous(), W
    if tag == "near":
        i = torch.from_numpy(blob["near/i"].astype(np.int64)); j = torch.from_numpy(blob["near/j"].astype(np.int64))
        near = (W[i] + W[j]) / 2 + torch.from_numpy(blob["near/noise"])
        return near.reshape(2, 10, 14, 10, D).permute(0, 4, 1, 2, 3).contiguous(), W
    if tag == "tiny":
        return torch.from_numpy(blob["tiny/z"]), W
    raise KeyError(tag)
