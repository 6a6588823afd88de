"""Token files / checkpoint names / rank sharding (host logic of SURVEY.md section 8(f) row 4; no GPU)."""
import os

import numpy as np
import pytest
import torch

from synthanatomy_b200.utils import tokens as tk


def test_subject_and_token_file_names(tmp_path):
    assert tk.subject_name("/data/ukb/sub-01_T1w.nii.gz") == "sub-01_T1w"
    assert tk.subject_name("sub-02.nii") == "sub-02"
    assert tk.subject_name("/x/y/sub-03") == "sub-03"
    p = tk.token_file_path(str(tmp_path), "/data/sub-01_T1w.nii.gz", level=0)
    assert p == os.path.join(str(tmp_path), "sub-01_T1w", "sub-01_T1w_quantization_0.npy")
    assert not os.path.exists(os.path.dirname(p))                     # makedirs is opt-in
    p1 = tk.token_file_path(str(tmp_path), "/data/sub-01_T1w.nii.gz", level=1, makedirs=True)
    assert p1.endswith("sub-01_T1w_quantization_1.npy") and os.path.isdir(os.path.dirname(p1))


def test_save_and_load_round_trip_uint16(tmp_path):
    g = torch.Generator().manual_seed(0)
    idx = torch.randint(0, 2048, (3, 10, 14, 10), generator=g)        # int64, as index_quantize returns
    names = [f"/data/sub-{i:02d}.nii.gz" for i in range(3)]
    paths = tk.save_token_volumes(idx, names, str(tmp_path))
    assert [os.path.basename(p) for p in paths] == [f"sub-{i:02d}_quantization_0.npy" for i in range(3)]
    for i, p in enumerate(paths):
        raw = np.load(p)
        assert raw.dtype == np.uint16 and raw.shape == (10, 14, 10)
        assert np.array_equal(tk.load_token_volume(p), idx[i].numpy().astype(np.uint16))
    with pytest.raises(ValueError):
        tk.save_token_volumes(torch.tensor([[[[70000]]]]), ["a.nii.gz"], str(tmp_path))
    with pytest.raises(ValueError):
        tk.save_token_volumes(idx, names[:2], str(tmp_path))
    assert tk.list_token_files(str(tmp_path)) == sorted(paths)


def test_checkpoint_path_resolution(tmp_path):
    d = str(tmp_path) + os.sep
    assert tk.checkpoint_path(d) is None
    for k in (2, 10, 9):
        torch.save({"k": k}, f"{d}checkpoint_epoch={k}.pt")
    assert tk.checkpoint_path(d).name == "checkpoint_epoch=10.pt"      # numeric, not lexicographic, order
    assert tk.checkpoint_path(d, epoch=9).name == "checkpoint_epoch=9.pt"
    with pytest.raises(FileNotFoundError):
        tk.checkpoint_path(d, epoch=3)
    with pytest.raises(RuntimeError):
        tk.checkpoint_path(d, which="best")
    torch.save({}, f"{d}checkpoint_key_metric=0.9.pt")
    assert tk.checkpoint_path(d, which="best").name == "checkpoint_key_metric=0.9.pt"


def test_token_batches_shard_disjointly_across_ranks(tmp_path):
    idx = torch.arange(10 * 2 * 3 * 2).reshape(10, 2, 3, 2) % 2048
    paths = tk.save_token_volumes(idx, [f"s{i}.nii.gz" for i in range(10)], str(tmp_path))
    seen = []
    for rank in range(2):
        it = tk.TokenBatches(paths, batch_size=2, rank=rank, world_size=2, shuffle_seed=7, pin_memory=False)
        assert len(it) == 2                                            # 5 files per rank, drop_last
        for b in it:
            assert b["quantization"].dtype == torch.uint16 and tuple(b["quantization"].shape) == (2, 2, 3, 2)
            seen.append((rank, tuple(b["filename_or_obj"])))
    files0 = {f for r, fs in seen if r == 0 for f in fs}
    files1 = {f for r, fs in seen if r == 1 for f in fs}
    assert len(files0) == 4 and len(files1) == 4 and not (files0 & files1)
    # same seed and epoch on every rank -> same permutation (rank 0 of a second loader reproduces the first epoch)
    again = [tuple(b["filename_or_obj"]) for b in tk.TokenBatches(paths, 2, 0, 2, shuffle_seed=7, pin_memory=False)]
    assert again == [fs for r, fs in seen if r == 0]
