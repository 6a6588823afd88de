"""Adversarial criteria and the engine-free adversarial iteration (host logic; stand-in networks on CPU)."""
import copy

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from synthanatomy_b200 import engines
from synthanatomy_b200.losses import AdversarialLoss, get_discriminator_loss, get_generator_loss


def test_adversarial_loss_formulas():
    g = torch.Generator().manual_seed(0)
    fake, real = torch.randn(2, 1, 3, 4, 5, generator=g), torch.randn(2, 1, 3, 4, 5, generator=g)
    ls_d = get_discriminator_loss({"discriminator_loss": "least_square"})
    ls_g = get_generator_loss({"generator_loss": "least_square"})
    assert ls_d.get_weight() == 0.005 and ls_g.get_weight() == 0.005
    torch.testing.assert_close(ls_d(fake, real), 0.005 * 0.5 * ((fake ** 2).mean() + ((real - 1) ** 2).mean()))
    torch.testing.assert_close(ls_g(fake), 0.005 * ((fake - 1) ** 2).mean())
    # the reference's names: "vanilla" is the relu hinge, "hinge" the softplus form
    van = AdversarialLoss("vanilla", True, weight=1.0)
    torch.testing.assert_close(van(fake, real), 0.5 * (F.relu(1 + fake).mean() + F.relu(1 - real).mean()))
    hin = AdversarialLoss("hinge", False, weight=2.0)
    torch.testing.assert_close(hin(fake), 2.0 * F.softplus(-fake).mean())
    assert set(ls_d.get_summaries()["scalar"]) == {"Loss-Adversarial_Discriminator-Reconstruction",
                                                   "Loss-Adversarial_Discriminator-Originals"}
    assert ls_g.set_weight(0.1) == 0.1
    with pytest.raises(ValueError):
        get_generator_loss({"generator_loss": "wasserstein"})
    with pytest.raises(ValueError):
        AdversarialLoss(reduction="none")


class _G(nn.Module):
    def __init__(self):
        super().__init__()
        self.body = nn.Conv3d(1, 4, 3, padding=1)
        self.last = nn.Conv3d(4, 1, 3, padding=1)

    def get_last_layer(self):
        return self.last.weight

    def forward(self, x):
        return {"reconstruction": [self.last(torch.tanh(self.body(x)))], "quantization_losses": [torch.zeros(())]}


def _recon(pred, y):
    return F.mse_loss(pred["reconstruction"][0], y) + pred["quantization_losses"][0]


@pytest.mark.parametrize("adaptive", [False, True])
def test_adversarial_iteration_matches_a_hand_written_step(adaptive):
    torch.manual_seed(1)
    G, D = _G(), nn.Sequential(nn.Conv3d(1, 2, 4, 2, 1), nn.LeakyReLU(0.2), nn.Conv3d(2, 1, 4, 1, 1))
    G2, D2 = copy.deepcopy(G), copy.deepcopy(D)
    x = torch.rand(2, 1, 8, 8, 8)
    lg, ld = get_generator_loss({"generator_loss": "least_square"}), get_discriminator_loss({"discriminator_loss": "hinge"})
    og, od = torch.optim.SGD(G.parameters(), lr=0.1), torch.optim.SGD(D.parameters(), lr=0.1)
    out = engines.adversarial_iteration(x, x, G, D, og, od, _recon, lg, ld, epoch=3, use_adversarial_adaptive_weight=adaptive,
                                        adaptive_adversarial_weight_threshold=0)
    # the same step written out by hand on the copies
    pred = G2(x)
    rec = _recon(pred, x)
    gl = 0.005 * ((D2(pred["reconstruction"][0]) - 1) ** 2).mean()
    if adaptive:
        a = torch.autograd.grad(rec, G2.last.weight, retain_graph=True)[0]
        b = torch.autograd.grad(gl, G2.last.weight, retain_graph=True)[0]
        w = torch.clamp(a.norm() / (b.norm() + 1e-4), 0.0, 1e4).detach()
    else:
        w = 1
    total = rec + gl * w
    gg = torch.autograd.grad(total, list(G2.parameters()))
    with torch.no_grad():
        for p, g_ in zip(G2.parameters(), gg):
            p -= 0.1 * g_
    fake = pred["reconstruction"][0].detach()
    dl = 0.005 * 0.5 * (F.softplus(D2(fake)).mean() + F.softplus(-D2(x)).mean()) * w
    dg = torch.autograd.grad(dl, list(D2.parameters()))
    with torch.no_grad():
        for p, g_ in zip(D2.parameters(), dg):
            p -= 0.1 * g_
    for p, q in zip(G.parameters(), G2.parameters()):
        torch.testing.assert_close(p, q)
    for p, q in zip(D.parameters(), D2.parameters()):       # in particular: no generator-pass gradient leaked into D's step
        torch.testing.assert_close(p, q)
    assert abs(out["loss"] - rec.item()) < 1e-6 and abs(out["g_loss"] - total.item()) < 1e-6 and abs(out["d_loss"] - dl.item()) < 1e-7
    assert set(out) == {"image", "label", "pred", "loss", "reals", "fakes", "g_loss", "d_loss"}


def test_adaptive_weight_threshold_and_disabled():
    G = _G()
    x = torch.rand(1, 1, 6, 6, 6)
    pred = G(x)
    rec = _recon(pred, x)
    gl = (pred["reconstruction"][0] ** 2).mean()
    assert engines.adaptive_adversarial_weight(G, rec, gl, 0, enabled=False) == 1
    assert engines.adaptive_adversarial_weight(G, rec, gl, 2, enabled=True, threshold=5, value=0.25) == 0.25
    w = engines.adaptive_adversarial_weight(G, rec, gl, 7, enabled=True, threshold=5, value=0.25)
    assert torch.is_tensor(w) and not w.requires_grad and 0.0 <= float(w) <= 1e4


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference/src"), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("criterion", ["vanilla", "hinge", "least_square"])
def test_adversarial_loss_against_the_imported_reference_class(criterion):
    """the unmodified AdversarialLoss (src/losses/adversarial/adversarial.py); only its TensorBoard enum import is stubbed"""
    import enum
    import sys
    import types
    hg = types.ModuleType("src.handlers.general")
    hg.TBSummaryTypes = enum.Enum("TBSummaryTypes", {"SCALAR": "scalar"})
    saved = {k: sys.modules.get(k) for k in ("src.handlers", "src.handlers.general")}
    sys.path.insert(0, "/root/reference")
    try:
        import src  # noqa: F401
        sys.modules["src.handlers"] = types.ModuleType("src.handlers")
        sys.modules["src.handlers.general"] = hg
        from src.losses.adversarial.adversarial import AdversarialLoss as Ref
    finally:
        sys.path.pop(0)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    g = torch.Generator().manual_seed(3)
    fake = torch.randn(2, 1, 3, 3, 3, generator=g).requires_grad_(True)
    real = torch.randn(2, 1, 3, 3, 3, generator=g)
    for is_d in (True, False):
        mine, ref = AdversarialLoss(criterion, is_d, weight=0.005), Ref(criterion=criterion, is_discriminator=is_d, weight=0.005)
        a = mine(fake, real if is_d else None)
        b = ref(fake, real if is_d else None)
        assert torch.equal(a, b)
        ga, = torch.autograd.grad(a, fake)
        gb, = torch.autograd.grad(b, fake)
        assert torch.equal(ga, gb)
