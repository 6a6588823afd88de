"""Host-side bookkeeping of ops.PackPlan (the per-pass multi-tensor packing of a stack's conv weights): which calls go to
the batched launch, which to the single one, and when a recorded entry may be handed out.  The library is replaced by a
counter -- no GPU, no arithmetic."""
import torch

from synthanatomy_b200 import ops


class _FakeLib:
    def __init__(self):
        self.single, self.multi = 0, []

    def sa_pack_weight(self, *a):
        self.single += 1
        return 0

    def sa_pack_weight_multi(self, arr, n, dt, st):
        self.multi.append(n)
        return 0


def _patch(monkeypatch):
    fake = _FakeLib()
    monkeypatch.setattr(ops, "lib", lambda: fake)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    monkeypatch.setattr(ops, "_p", lambda t: None if t is None else t.data_ptr())     # (the real one refuses CPU tensors)
    return fake


def test_first_pass_records_later_passes_batch(monkeypatch):
    fake = _patch(monkeypatch)
    w1, w2 = torch.zeros(4, 3, 2, 2, 2), torch.zeros(5, 4, 1, 1, 1)
    plan = ops.PackPlan()
    with ops.pack_plan(plan):
        a1 = ops.pack_weight(w1, False, torch.float32)
        a1t = ops.pack_weight(w1, True, torch.float32)
    assert (fake.single, fake.multi) == (2, []) and tuple(a1.shape) == (8, 4, 3) and tuple(a1t.shape) == (8, 3, 4)
    with ops.pack_plan(plan):                               # second pass: both forms refreshed by ONE batched call
        assert ops.pack_weight(w1, False, torch.float32) is a1
        assert ops.pack_weight(w1, True, torch.float32) is a1t
        a2 = ops.pack_weight(w2, False, torch.float32)      # a weight the plan has not seen: packed singly, recorded
    assert (fake.single, fake.multi) == (3, [2])
    with ops.pack_plan(plan, begin=False):                  # the backward pass of that forward: no refresh, same entries
        assert ops.pack_weight(w1, True, torch.float32) is a1t and ops.pack_weight(w2, False, torch.float32) is a2
    assert (fake.single, fake.multi) == (3, [2])
    with ops.pack_plan(plan):
        assert ops.pack_weight(w2, False, torch.float32) is a2
    assert (fake.single, fake.multi) == (3, [2, 3])


def test_entries_are_valid_for_the_current_pass_only(monkeypatch):
    fake = _patch(monkeypatch)
    w = torch.zeros(4, 3, 1, 1, 1)
    plan = ops.PackPlan()
    with ops.pack_plan(plan):
        a = ops.pack_weight(w, False, torch.float32)
    plan.epoch += 1                                         # (what a begin() of another user of the plan would do)
    with ops.pack_plan(plan, begin=False):
        b = ops.pack_weight(w, False, torch.float32)        # stale stamp: packed again, not handed out
    assert b is not a and fake.single == 2
    # dtypes are separate entries and separate batched calls
    with ops.pack_plan(plan):
        ops.pack_weight(w, False, torch.bfloat16)
    with ops.pack_plan(plan):
        pass
    assert sorted(fake.multi[-2:]) == [1, 1]


def test_without_a_plan_every_call_packs(monkeypatch):
    fake = _patch(monkeypatch)
    w = torch.zeros(4, 3, 1, 1, 1)
    assert ops.current_pack_plan() is None
    a, b = ops.pack_weight(w, False, torch.float32), ops.pack_weight(w, False, torch.float32)
    assert a is not b and fake.single == 2 and fake.multi == []


def test_plan_does_not_grow_without_bound(monkeypatch):
    _patch(monkeypatch)
    plan = ops.PackPlan()
    keep = []
    with ops.pack_plan(plan):
        for _ in range(1100):                               # weights that move every pass (dtype-converted copies)
            w = torch.zeros(2, 2, 1, 1, 1)
            keep.append(w)
            ops.pack_weight(w, False, torch.float32)
    assert len(plan.entries) <= 1024
