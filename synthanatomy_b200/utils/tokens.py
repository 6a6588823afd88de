"""Token files and device-side batch preparation: the data that travels between the VQ-VAE and the Performer.

On disk (written by the extraction mode, read by the transformer's training mode):
    <outputs_directory>/<subject>/<subject>_quantization_<level>.npy      uint16, one latent index grid [D, H, W] per subject
which is what NpySaver (/root/reference/src/handlers/general.py:491-590, attached at /root/reference/run_vqvae.py:484-498)
produces through MONAI's create_file_basename (separate folder per subject, ".nii.gz" stripped, "_<postfix>" appended),
and checkpoints live at <checkpoint_directory>checkpoint_epoch=<K>.pt (/root/reference/src/utils/general.py:75-168).

On the device: `prepare_batch_device` uploads the grid in its stored type (2 bytes per token) and forms the int64 input /
target sequences with one gather kernel (sa_tokens_prepare) instead of the host-side reshape / gather / pad / widen chain of
/root/reference/src/utils/transformer.py:259-282 followed by two 8-byte-per-token copies; `sequence_to_grid` is the inverse
used after sampling (/root/reference/src/inferer/transformer.py:63-71 + the ordering's reverted index sequence).
"""
from __future__ import annotations

import glob
import os
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

TOKEN_DTYPE = np.uint16


# ------------------------------------------------------------------------------------------------ file names
def subject_name(filename_or_obj: str) -> str:
    """'/data/sub-01_T1w.nii.gz' -> 'sub-01_T1w' (one extension stripped, two for '.gz')"""
    name = os.path.basename(str(filename_or_obj))
    name, ext = os.path.splitext(name)
    if ext == ".gz":
        name, _ = os.path.splitext(name)
    return name


def token_file_path(output_dir: str, filename_or_obj: str, level: int = 0, makedirs: bool = False) -> str:
    """path of the token file of one subject: <output_dir>/<subject>/<subject>_quantization_<level>.npy"""
    subject = subject_name(filename_or_obj)
    folder = os.path.join(output_dir, subject)
    if makedirs:
        os.makedirs(folder, exist_ok=True)
    return os.path.normpath(os.path.join(folder, f"{subject}_quantization_{level}.npy"))


def checkpoint_path(checkpoint_directory: str, epoch: int = -1, which: str = "recent") -> Optional[Path]:
    """<checkpoint_directory>checkpoint_epoch=<K>.pt for a given epoch, the most recent one (epoch == -1,
    which == 'recent') or the single key-metric checkpoint (which == 'best'); None if nothing is there.
    (the directory string is used as a prefix, as the reference does: it is expected to end with a separator)"""
    if epoch > 0:
        p = Path(f"{checkpoint_directory}checkpoint_epoch={epoch}.pt")
        if not p.exists():
            raise FileNotFoundError(f"Checkpoint '{p.as_posix()}' is not found.")
        return p
    if which == "best":
        found = glob.glob(checkpoint_directory + "checkpoint_key_metric*.pt")
        if len(found) != 1:
            raise RuntimeError(f"Should only be one best metric checkpoint, found {found}")
        return Path(found[0])
    epochs = sorted(int(os.path.basename(e).split("_")[-1].split("=")[-1].split(".")[0])
                    for e in glob.glob(checkpoint_directory + "*checkpoint_epoch*.pt"))
    return Path(f"{checkpoint_directory}checkpoint_epoch={epochs[-1]}.pt") if epochs else None


# ------------------------------------------------------------------------------------------------ token files
def _to_token_array(indices) -> np.ndarray:
    if torch.is_tensor(indices):
        if indices.is_cuda and indices.dtype == torch.int64:
            from .. import pf_ops                      # narrow on the device: 2 bytes per token cross PCIe instead of 8
            indices = pf_ops.tokens_narrow(indices.contiguous())
        indices = indices.detach().cpu()
        if indices.dtype == torch.uint16:
            return indices.numpy()
        indices = indices.numpy()
    arr = np.asarray(indices)
    if arr.size and (arr.min() < 0 or arr.max() > np.iinfo(TOKEN_DTYPE).max):
        raise ValueError("token index outside [0, 65535] cannot be stored as uint16")
    return arr.astype(TOKEN_DTYPE)


def save_token_volumes(indices, filenames: Sequence[str], output_dir: str, level: int = 0) -> List[str]:
    """Store a batch of latent index grids [B, D, H, W] (as returned by index_quantize()[level]) one .npy per subject."""
    arr = _to_token_array(indices)
    if arr.shape[0] != len(filenames):
        raise ValueError(f"{arr.shape[0]} grids but {len(filenames)} file names")
    paths = []
    for grid, name in zip(arr, filenames):
        path = token_file_path(output_dir, name, level, makedirs=True)
        np.save(file=path, arr=np.ascontiguousarray(grid))
        paths.append(path)
    return paths


def load_token_volume(path: str) -> np.ndarray:
    arr = np.load(path)
    if arr.dtype != TOKEN_DTYPE:
        if not np.issubdtype(arr.dtype, np.integer):
            raise TypeError(f"{path}: expected integer tokens, found {arr.dtype}")
        arr = arr.astype(TOKEN_DTYPE)
    return arr


def list_token_files(subjects: str) -> List[str]:
    """a folder (searched recursively for *_quantization_*.npy, sorted) or a .csv / .tsv with a 'subject' column"""
    if os.path.isdir(subjects):
        return sorted(glob.glob(os.path.join(subjects, "**", "*.npy"), recursive=True))
    sep = "\t" if subjects.endswith(".tsv") else ","
    with open(subjects) as f:
        header = f.readline().rstrip("\n").split(sep)
        col = header.index("subject")
        return [line.rstrip("\n").split(sep)[col] for line in f if line.strip()]


class TokenBatches:
    """Iterates pinned uint16 batches {"quantization": [B, D, H, W], "filename_or_obj": [...]} over token files;
    rank-sharded (every rank sees a disjoint, equally sized slice, as DistributedSampler with drop_last would give)."""

    def __init__(self, files: Iterable[str], batch_size: int, rank: int = 0, world_size: int = 1, shuffle_seed: Optional[int] = None,
                 drop_last: bool = True, pin_memory: Optional[bool] = None):
        self.files = list(files)
        self.batch_size, self.rank, self.world_size = int(batch_size), int(rank), int(world_size)
        self.shuffle_seed, self.drop_last = shuffle_seed, drop_last
        self.pin_memory = torch.cuda.is_available() if pin_memory is None else pin_memory
        self.epoch = 0

    def _order(self) -> List[int]:
        idx = list(range(len(self.files)))
        if self.shuffle_seed is not None:
            rng = np.random.default_rng(self.shuffle_seed + self.epoch)
            rng.shuffle(idx)
        per_rank = len(idx) // self.world_size
        return idx[self.rank * per_rank:(self.rank + 1) * per_rank]

    def __len__(self) -> int:
        n = len(self.files) // self.world_size
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def __iter__(self):
        order = self._order()
        self.epoch += 1
        for s in range(0, len(order), self.batch_size):
            chunk = order[s:s + self.batch_size]
            if len(chunk) < self.batch_size and self.drop_last:
                return
            grids = np.stack([load_token_volume(self.files[i]) for i in chunk])
            t = torch.from_numpy(grids)
            if self.pin_memory:
                t = t.pin_memory()
            yield {"quantization": t, "filename_or_obj": [self.files[i] for i in chunk]}


# ------------------------------------------------------------------------------------------------ device side
class DeviceOrdering:
    """the ordering's index sequences, resident on the device (built once per (ordering, device))"""

    def __init__(self, ordering, device):
        self.device = torch.device(device)
        self.index_sequence = torch.as_tensor(np.asarray(ordering.get_sequence_ordering()), dtype=torch.int64).to(self.device)
        self.revert_sequence = torch.as_tensor(np.asarray(ordering.get_revert_sequence_ordering()), dtype=torch.int64).to(self.device)


def prepare_batch_device(batch: Dict, order: DeviceOrdering, vocab_size: int, conditionings=None, non_blocking: bool = True):
    """Device-side prepare_batch: same ((x_input, conditioned), x_target) as utils.transformer.prepare_batch, with the
    token grid crossing PCIe in its stored type and ONE gather launch forming both int64 sequences."""
    from .. import pf_ops
    grid = batch["quantization"]
    if not torch.is_tensor(grid):
        grid = torch.from_numpy(np.ascontiguousarray(grid))
    if grid.dtype not in (torch.uint16, torch.int16, torch.int32, torch.int64):
        grid = grid.long()
    grid = grid.to(order.device, non_blocking=non_blocking).contiguous()
    x_input, x_target = pf_ops.tokens_prepare(grid, order.index_sequence, vocab_size)
    conditioned = None
    if conditionings:
        conditioned = []
        for label in conditionings:
            c = batch[label]
            if c.dim() == 1:
                c = c[..., None]
            conditioned.append(c.long().to(order.device, non_blocking=non_blocking))
    return (x_input, conditioned), x_target


def sequence_to_grid(sequence: torch.Tensor, order: DeviceOrdering, dimensions: Sequence[int]) -> torch.Tensor:
    """sampled token sequence [B, N] (prefix already stripped) -> latent index grid: the tail of TransformerBase.sample
    (/root/reference/src/networks/transformers/transformer.py:95-99: x[:, revert] -> reshape(B, *ordering.dimensions) ->
    squeeze(1)) as one device gather, int64 out, ready for decode_samples"""
    from .. import pf_ops
    seq = sequence.to(order.device).contiguous()
    if seq.dtype not in (torch.uint16, torch.int16, torch.int32, torch.int64):
        seq = seq.long()
    out = pf_ops.tokens_gather(seq, order.revert_sequence)
    return torch.squeeze(out.reshape(seq.shape[0], *dimensions), 1)
