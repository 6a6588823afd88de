"""Host-side cache of the local heads' (cos, sin) tables (performer._rot_table): one table per frequency buffer, sequence
length and head dimension; recomputed when the buffer is written in place (load_state_dict) or replaced.  The library call
is replaced by a counter -- no GPU, no arithmetic."""
import torch

from synthanatomy_b200 import pf_ops
from synthanatomy_b200.networks.transformers import performer as pm


def _patch(monkeypatch):
    calls = []

    def fake_table(inv_freq, seq, dim_head):
        calls.append((inv_freq.data_ptr(), seq, dim_head))
        return torch.zeros(seq, dim_head // 2, 2)

    monkeypatch.setattr(pf_ops, "rotary_table", fake_table)
    monkeypatch.setattr(pm, "_ROT_TABLES", {})
    return calls


def test_one_table_per_buffer_length_and_head_dimension(monkeypatch):
    calls = _patch(monkeypatch)
    f1, f2 = torch.rand(32), torch.rand(32)
    t = pm._rot_table(f1, 100, 64)
    assert pm._rot_table(f1, 100, 64) is t and len(calls) == 1          # every layer's backward of every step hits
    pm._rot_table(f2, 100, 64)                                           # another layer's buffer
    pm._rot_table(f1, 101, 64)                                           # another sequence length
    assert len(calls) == 3
    assert pm._rot_table(f2, 100, 64).shape == (100, 32, 2) and len(calls) == 3


def test_in_place_write_of_the_buffer_invalidates_its_table(monkeypatch):
    calls = _patch(monkeypatch)
    f = torch.rand(32)
    t = pm._rot_table(f, 50, 64)
    f.copy_(torch.rand(32))                                              # what load_state_dict does
    assert pm._rot_table(f, 50, 64) is not t and len(calls) == 2


def test_entries_keep_their_buffer_alive_and_the_cache_is_bounded(monkeypatch):
    calls = _patch(monkeypatch)
    f = torch.rand(32)
    ptr = f.data_ptr()
    pm._rot_table(f, 10, 64)
    del f                                                                # the address cannot be handed to another tensor
    assert any(k[1] == ptr for k in pm._ROT_TABLES)
    for n in range(400):
        pm._rot_table(torch.rand(32), 10 + n, 64)
    assert len(pm._ROT_TABLES) <= 256 and len(calls) == 401
