"""The discriminator oracle against the golden vectors produced by the unmodified reference class
(oracle/make_golden_discriminator.py), and the drop-in module's parameter tree against the same file (no GPU)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import discriminator_oracle as do

GOLD = os.path.join(os.path.dirname(__file__), "golden", "discriminator.npz")


def _load():
    g = np.load(GOLD)
    state = {k[5:]: torch.from_numpy(g[k].copy()) for k in g.files if k.startswith("init/")}
    return g, state


def test_oracle_reproduces_reference_forward_backward_and_buffers():
    g, state = _load()
    for k, v in state.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    x = torch.from_numpy(g["x"].copy()).requires_grad_(True)
    out = do.forward(state, x, training=True)
    assert torch.equal(out.detach(), torch.from_numpy(g["out"]))            # same torch CPU kernels: exact
    loss = ((out - 1.0) ** 2).mean()
    assert abs(loss.item() - float(g["loss"])) <= 1e-7
    loss.backward()
    torch.testing.assert_close(x.grad, torch.from_numpy(g["dx"]), rtol=1e-6, atol=1e-9)
    for k in g.files:
        if k.startswith("grad/"):
            torch.testing.assert_close(state[k[5:]].grad, torch.from_numpy(g[k]), rtol=1e-5, atol=1e-8, msg=k)
        if k.startswith("after/"):
            torch.testing.assert_close(state[k[6:]].detach(), torch.from_numpy(g[k]), rtol=1e-6, atol=1e-8, msg=k)
    with torch.no_grad():
        out_eval = do.forward(state, x, training=False)
    torch.testing.assert_close(out_eval, torch.from_numpy(g["out_eval"]), rtol=1e-6, atol=1e-7)


def test_layer_plan_matches_patchgan_layout():
    _, state = _load()
    assert do.layer_plan(state) == [(0, 2, False, True), (2, 2, True, True), (5, 2, True, True), (8, 1, True, True),
                                    (11, 1, False, False)]


def test_dropin_module_has_the_reference_parameter_tree_and_init_stream():
    from synthanatomy_b200.networks.discriminator import B200Discriminator
    g, state = _load()
    torch.manual_seed(11)
    net = B200Discriminator(input_nc=1, ndf=8, n_layers=3)
    sd = net.state_dict()
    assert list(sd.keys()) == list(state.keys())
    for k in sd:
        assert torch.equal(sd[k], state[k].detach()), k                  # same modules, same order => same random stream
    with pytest.raises(RuntimeError):
        net(torch.rand(1, 1, 32, 32, 32))                                   # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        net.main(torch.rand(1, 1, 32, 32, 32))                              # containers are not executable


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present (GPU box)")
def test_oracle_against_the_imported_reference_on_a_second_configuration():
    sys.path.insert(0, "/root/reference")
    try:
        from src.networks.discriminator.baseline import BaselineDiscriminator
    finally:
        sys.path.pop(0)
    torch.manual_seed(3)
    ref = BaselineDiscriminator(input_nc=1, ndf=4, n_layers=2).train()
    state = {k: v.clone() for k, v in ref.state_dict().items()}
    x = torch.rand(1, 1, 24, 24, 24)
    with torch.no_grad():
        assert torch.equal(do.forward(state, x, training=True), ref(x))
    for k, v in ref.state_dict().items():
        torch.testing.assert_close(state[k], v, rtol=1e-6, atol=1e-8, msg=k)
