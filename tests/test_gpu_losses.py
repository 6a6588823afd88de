"""the spectral (Jukebox) reconstruction loss on the GPU -- dense DFT-matrix products on the tensor
cores in bf16x3 arithmetic -- against the golden value / gradient of the unmodified reference class
(tests/golden/jukebox.npz, made by oracle/make_golden_losses.py) and against the pinned CPU oracle at larger shapes.
Tolerance: 1e-4 relative on the loss, 1e-4 of max |grad| on the gradient."""
import os

import numpy as np
import pytest
import torch

from oracle import losses_oracle as lo

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "jukebox.npz")


def test_swap_outer_inner_kernel():
    from synthanatomy_b200 import losses
    g = torch.Generator().manual_seed(0)
    for b, A, M, Cc in ((3, 5, 2, 7), (2, 33, 4, 65), (1, 160, 6, 224)):
        src = torch.randn(b, A, M, Cc, generator=g)
        got = losses._swap(src.cuda(), b, A, M, Cc).cpu().view(b, Cc, M, A)
        assert torch.equal(got, src.permute(0, 3, 2, 1).contiguous())


def test_jukebox_loss_matches_reference_golden_value_and_gradient():
    from synthanatomy_b200.losses import JukeboxLoss
    g = np.load(GOLD)
    y, q = torch.from_numpy(g["y"]).cuda(), torch.from_numpy(g["q"]).cuda()
    for name, kw, factor in (("default", {}, 1.0), ("no_pixel_f2", {"include_pixel_loss": False}, 2.0)):
        pred = torch.from_numpy(g["pred"].copy()).cuda().requires_grad_(True)
        crit = JukeboxLoss(dimensions=3, **kw)
        assert crit.set_fft_factor(factor) == factor
        loss = crit({"reconstruction": [pred], "quantization_losses": [q]}, y)
        want = float(g[f"{name}/loss"])
        assert abs(float(loss) - want) <= 1e-4 * abs(want), (name, float(loss), want)
        loss.backward()
        gref = torch.from_numpy(g[f"{name}/grad"])
        err = float((pred.grad.cpu() - gref).abs().max()) / float(gref.abs().max())
        assert err <= 1e-4, (name, err)
        assert "Loss-Spectral-Reconstruction" in crit.get_summaries()["scalar"]


@pytest.mark.parametrize("shape", [(2, 1, 32, 48, 40), (1, 1, 80, 112, 80)])
def test_spectral_loss_against_oracle_at_volume_shapes(shape):
    """non-power-of-two axes with the factors of the README volume (2^5 x 5, 2^5 x 7); the second case is a level-1-sized
    volume (the tensor-core path of every product is exercised: 716 800 voxels)"""
    from synthanatomy_b200 import ops
    from synthanatomy_b200.losses import spectral_loss
    g = torch.Generator().manual_seed(sum(shape))
    y = torch.rand(shape, generator=g)
    pred = (y + 0.3 * torch.randn(shape, generator=g)).requires_grad_(True)
    want = lo.jukebox_loss(pred, y, (), include_pixel_loss=False)
    want.backward()
    p = pred.detach().cuda().requires_grad_(True)
    loss = spectral_loss(p, y.cuda())
    assert ops.last_path() == 2, "the DFT products did not run on the tcgen05 GEMM"
    assert abs(float(loss) - float(want)) <= 1e-4 * abs(float(want)), (float(loss), float(want))
    loss.backward()
    err = float((p.grad.cpu() - pred.grad).abs().max()) / float(pred.grad.abs().max())
    assert err <= 1e-4, err


def test_jukebox_loss_rejects_what_it_does_not_implement():
    from synthanatomy_b200.losses import JukeboxLoss
    with pytest.raises(NotImplementedError):
        JukeboxLoss(dimensions=2)
    with pytest.raises(NotImplementedError):
        JukeboxLoss(dimensions=3, fft_kwargs={"s": None, "dim": (2, 3, 4), "norm": "ortho"})
