"""Checkpoint files in the layout the reference's runs read and write.

MONAI's ``CheckpointSaver`` / ``CheckpointLoader`` (attached at /root/reference/run_vqvae.py:312-361 and
run_transformer.py:234-264) store one ``torch.save``d dict per file, ``{name: obj.state_dict()}`` for
``network, optimizer, lr_scheduler, trainer`` (+ ``d_network, d_optimizer, d_lr_scheduler`` with the adversarial
component), DistributedDataParallel wrappers unwrapped, as ``<checkpoint_directory>checkpoint_epoch=<K>.pt`` with
``n_saved=1`` (the previous epoch's file is removed) and ``checkpoint_key_metric=<v>.pt`` for the best validation score.
``trainer`` is the Ignite engine state: ``{"epoch_length", "max_epochs", "iteration"}``.

These helpers write and read exactly that, so a run can resume from -- and be resumed by -- the reference's own
entry points (the drop-in modules keep the reference's ``state_dict`` keys).  File discovery: ``tokens.checkpoint_path``.
"""
from __future__ import annotations

import glob
import os
from collections import OrderedDict
from typing import Dict, Optional

import torch


class TrainerState:
    """the part of an Ignite ``Engine`` that its ``state_dict`` carries"""

    def __init__(self, epoch_length: Optional[int] = None, max_epochs: Optional[int] = None, iteration: int = 0):
        self.epoch_length, self.max_epochs, self.iteration = epoch_length, max_epochs, iteration

    @property
    def epoch(self) -> int:
        return 0 if not self.epoch_length else self.iteration // self.epoch_length

    def state_dict(self) -> "OrderedDict[str, int]":
        return OrderedDict(epoch_length=self.epoch_length, max_epochs=self.max_epochs, iteration=self.iteration)

    def load_state_dict(self, sd) -> None:
        self.epoch_length, self.max_epochs = sd["epoch_length"], sd["max_epochs"]
        if "iteration" in sd:
            self.iteration = sd["iteration"]
        else:                                       # Ignite accepts either "iteration" or "epoch"
            self.iteration = sd["epoch"] * (self.epoch_length or 0)


def _unwrap(obj):
    return obj.module if isinstance(obj, (torch.nn.parallel.DistributedDataParallel, torch.nn.DataParallel)) else obj


def save_checkpoint(objects: Dict[str, object], checkpoint_directory: str, epoch: int, n_saved: Optional[int] = 1,
                    key_metric: Optional[float] = None) -> str:
    """Write ``{name: state_dict}``; returns the path.  ``checkpoint_directory`` is a prefix ending in a separator, as in
    the reference.  With ``key_metric`` the file is the single best-metric checkpoint instead of an epoch checkpoint."""
    os.makedirs(os.path.dirname(checkpoint_directory) or ".", exist_ok=True)
    payload = {name: _unwrap(obj).state_dict() for name, obj in objects.items()}
    if key_metric is not None:
        for old in glob.glob(checkpoint_directory + "checkpoint_key_metric=*.pt"):
            os.remove(old)
        path = f"{checkpoint_directory}checkpoint_key_metric={key_metric:.4f}.pt"
    else:
        path = f"{checkpoint_directory}checkpoint_epoch={epoch}.pt"
    tmp = path + ".tmp"
    torch.save(payload, tmp)
    os.replace(tmp, path)                            # a killed run never leaves a truncated checkpoint behind
    if key_metric is None and n_saved is not None:
        found = sorted(glob.glob(checkpoint_directory + "checkpoint_epoch=*.pt"),
                       key=lambda p: int(os.path.basename(p).split("=")[-1].split(".")[0]))
        for old in found[:-n_saved] if n_saved > 0 else found:
            if old != path:
                os.remove(old)
    return path


def load_checkpoint(path: str, objects: Dict[str, object], map_location=None, strict: bool = True) -> Dict:
    """Restore every object named in ``objects`` from the file (MONAI's CheckpointLoader: a name missing from the file is
    an error when ``strict``); returns the raw dict."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    for name, obj in objects.items():
        if name not in ckpt:
            if strict:
                raise KeyError(f"checkpoint '{path}' has no entry '{name}' (found {sorted(ckpt)})")
            continue
        _unwrap(obj).load_state_dict(ckpt[name])
    return ckpt


def save_model_state_dict(network, checkpoint_directory: str, epoch: int) -> str:
    """the bare ``state_dict`` the training entry points leave at the end of a run (run_vqvae.py:389-392; like the
    reference it is taken from the object as passed, so a DistributedDataParallel wrapper contributes its ``module.`` prefix)"""
    path = f"{checkpoint_directory}model_state_dict_epoch={epoch}.pt"
    torch.save(network.state_dict(), path)
    return path
