"""Checkpoint files in the reference's layout (MONAI CheckpointSaver / CheckpointLoader dicts); no GPU."""
import os

import pytest
import torch

from synthanatomy_b200.optim import Adam
from synthanatomy_b200.networks.discriminator import B200Discriminator
from synthanatomy_b200.utils import checkpoints as ck
from synthanatomy_b200.utils import tokens as tk


def _objects(seed):
    torch.manual_seed(seed)
    net = B200Discriminator(ndf=4, n_layers=2)
    opt = Adam(net.parameters(), lr=5e-4)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.99)
    return {"network": net, "optimizer": opt, "lr_scheduler": sched, "trainer": ck.TrainerState(100, 50, 300)}


def test_save_writes_the_reference_layout_and_keeps_one_file(tmp_path):
    d = str(tmp_path) + os.sep
    objs = _objects(0)
    p2 = ck.save_checkpoint(objs, d, epoch=2)
    assert os.path.basename(p2) == "checkpoint_epoch=2.pt"
    p3 = ck.save_checkpoint(objs, d, epoch=3)
    assert sorted(os.listdir(d)) == ["checkpoint_epoch=3.pt"]            # n_saved = 1
    raw = torch.load(p3, weights_only=False)
    assert set(raw) == {"network", "optimizer", "lr_scheduler", "trainer"}
    assert list(raw["network"].keys()) == list(objs["network"].state_dict().keys())
    assert dict(raw["trainer"]) == {"epoch_length": 100, "max_epochs": 50, "iteration": 300}
    assert tk.checkpoint_path(d).name == "checkpoint_epoch=3.pt"
    best = ck.save_checkpoint(objs, d, epoch=3, key_metric=0.91237)
    assert os.path.basename(best) == "checkpoint_key_metric=0.9124.pt"
    ck.save_checkpoint(objs, d, epoch=4, key_metric=0.95)
    assert tk.checkpoint_path(d, which="best").name == "checkpoint_key_metric=0.9500.pt"
    assert os.path.basename(ck.save_model_state_dict(objs["network"], d, 4)) == "model_state_dict_epoch=4.pt"


def test_load_restores_a_checkpoint_written_with_stock_torch_objects(tmp_path):
    """what a reference run leaves behind: the reference class's state_dict + torch.optim.Adam (tensor step counters)"""
    import torch.nn as nn
    torch.manual_seed(1)
    src = B200Discriminator(ndf=4, n_layers=2)            # same state_dict keys as the reference class (tested elsewhere)
    params = list(src.parameters())
    opt = torch.optim.Adam(params, lr=5e-4)
    for p in params:
        p.grad = torch.randn_like(p)
    opt.step()
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.99)
    sched.step()
    path = str(tmp_path / "checkpoint_epoch=7.pt")
    torch.save({"network": src.state_dict(), "optimizer": opt.state_dict(), "lr_scheduler": sched.state_dict(),
                "trainer": {"epoch_length": 10, "max_epochs": 20, "epoch": 7}, "d_network": nn.Linear(2, 2).state_dict()}, path)
    objs = _objects(2)
    ck.load_checkpoint(path, objs, map_location="cpu")
    for k, v in objs["network"].state_dict().items():
        assert torch.equal(v, src.state_dict()[k]), k
    st = objs["optimizer"].state[next(iter(objs["network"].parameters()))]
    assert torch.equal(st["exp_avg"], opt.state[params[0]]["exp_avg"]) and float(st["step"]) == 1.0
    assert objs["optimizer"].param_groups[0]["lr"] == pytest.approx(opt.param_groups[0]["lr"])
    assert objs["trainer"].iteration == 70 and objs["trainer"].epoch == 7
    with pytest.raises(KeyError):
        ck.load_checkpoint(path, {"g_network": objs["network"]})
    assert "d_network" in ck.load_checkpoint(path, {"g_network": objs["network"]}, strict=False)


def test_ddp_style_wrappers_are_unwrapped():
    objs = _objects(3)
    wrapped = torch.nn.DataParallel(objs["network"])
    assert list(ck._unwrap(wrapped).state_dict()) == list(objs["network"].state_dict())


def test_optimizer_state_round_trips_into_torch_adam_and_steps():
    """a run saved with this library's Adam is resumed by the reference's entry points with torch.optim.Adam
    (run_vqvae.py:82, :312-344): the param_groups must carry every key torch's step() reads"""
    torch.manual_seed(5)
    src = torch.nn.Linear(3, 2)
    mine = Adam(src.parameters(), lr=5e-4)
    for p in src.parameters():            # the state a few steps of the CUDA kernel would have left
        mine.state[p] = {"step": torch.tensor(3.0), "exp_avg": torch.randn_like(p), "exp_avg_sq": torch.rand_like(p)}
    sd = mine.state_dict()
    ref_keys = set(torch.optim.Adam(torch.nn.Linear(1, 1).parameters()).param_groups[0])
    assert ref_keys <= set(sd["param_groups"][0]), ref_keys - set(sd["param_groups"][0])
    dst = torch.nn.Linear(3, 2)
    dst.load_state_dict(src.state_dict())
    theirs = torch.optim.Adam(dst.parameters(), lr=1.0)
    theirs.load_state_dict(sd)
    before = [p.detach().clone() for p in dst.parameters()]
    for p in dst.parameters():
        p.grad = torch.ones_like(p)
    theirs.step()                                                        # raised KeyError('weight_decay') before
    assert theirs.param_groups[0]["lr"] == 5e-4
    assert all(float(theirs.state[p]["step"]) == 4.0 for p in dst.parameters())
    assert all(not torch.equal(a, p) for a, p in zip(before, dst.parameters()))
    # and back: a torch checkpoint with a setting the kernel does not implement is refused, not ignored
    wd = torch.optim.Adam(torch.nn.Linear(3, 2).parameters(), lr=1e-3, weight_decay=0.1)
    with pytest.raises(NotImplementedError):
        Adam(torch.nn.Linear(3, 2).parameters()).load_state_dict(wd.state_dict())
    with pytest.raises(NotImplementedError):
        Adam(torch.nn.Linear(3, 2).parameters(), amsgrad=True)
