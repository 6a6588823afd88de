"""B200 discriminator (conv kernels + BatchNorm3d / LeakyReLU kernels, hand-scheduled backward) against the CPU oracle
and the golden vectors of the unmodified reference class.  fp32 path: 1e-4; bf16 path: tracks within bf16 rounding."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import discriminator_oracle as do

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "discriminator.npz")


def _rel(a, b):
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-12)


@pytest.mark.parametrize("rows,C,offset", [(1000, 64, 0.0), (4099, 130, 50.0), (37, 8, -3.0)])
def test_batchnorm_lrelu_kernels_match_torch(rows, C, offset):
    from synthanatomy_b200 import ops
    g = torch.Generator().manual_seed(rows + C)
    x = torch.randn(rows, C, generator=g) * 2.0 + offset               # offset >> std: two-pass variance must hold up
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    gy = torch.randn(rows, C, generator=g)
    rm, rv = torch.zeros(C), torch.ones(C)
    # reference in float64: torch's own fp32 CPU BatchNorm backward loses half its digits at offset / std = 25
    xr = x.double().requires_grad_(True)
    gr, br = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rm_ref, rv_ref = rm.double(), rv.double()
    y_ref = F.leaky_relu(F.batch_norm(xr.t()[None], rm_ref, rv_ref, gr, br, True, 0.1, 1e-5)[0].t(), 0.2)
    y_ref.backward(gy.double())
    rm_ref, rv_ref = rm_ref.float(), rv_ref.float()
    xd = x.cuda()
    rmd, rvd = rm.cuda(), rv.cuda()
    mean, rstd = ops.bn_stats(xd, 1e-5, 0.1, rmd, rvd)
    y = ops.bn_lrelu_fwd(xd, mean, rstd, gamma.cuda(), beta.cuda(), 0.2)
    torch.testing.assert_close(rmd.cpu(), rm_ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rvd.cpu(), rv_ref, rtol=1e-4, atol=1e-6)
    assert _rel(y.cpu().double(), y_ref.detach()) <= 1e-4
    dx, dgamma, dbeta = ops.bn_lrelu_bwd(gy.cuda(), xd, y, mean, rstd, gamma.cuda(), 0.2)
    assert _rel(dx.cpu().double(), xr.grad) <= 2e-4
    assert _rel(dgamma.cpu().double(), gr.grad) <= 1e-4 and _rel(dbeta.cpu().double(), br.grad) <= 1e-4
    # eval-mode statistics
    m2, r2 = ops.bn_eval_stats(rmd, rvd, 1e-5)
    torch.testing.assert_close(r2.cpu(), 1.0 / torch.sqrt(rv_ref + 1e-5), rtol=1e-5, atol=0)
    # standalone LeakyReLU pair
    z = xd.clone()
    ops.lrelu_fwd_(z, 0.2)
    assert torch.equal(z.cpu(), F.leaky_relu(x, 0.2))
    gz = gy.cuda().clone()
    ops.lrelu_bwd_(gz, z, 0.2)
    assert torch.equal(gz.cpu(), gy * torch.where(x > 0, 1.0, 0.2))


def test_discriminator_matches_reference_golden_vectors_fp32():
    from synthanatomy_b200.networks.discriminator import B200Discriminator
    g = np.load(GOLD)
    torch.manual_seed(11)
    net = B200Discriminator(input_nc=1, ndf=8, n_layers=3).cuda().train()
    for k, v in net.state_dict().items():
        assert torch.equal(v.cpu(), torch.from_numpy(g["init/" + k])), k
    x = torch.from_numpy(g["x"].copy()).cuda().requires_grad_(True)
    out = net(x)
    assert tuple(out.shape) == (2, 1, 2, 2, 2)
    assert _rel(out.detach().cpu(), torch.from_numpy(g["out"])) <= 1e-4
    assert _rel(out.detach().cpu().double(), torch.from_numpy(g["f64/out"])) <= 1e-4
    loss = ((out - 1.0) ** 2).mean()
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= 1e-4 * max(1.0, float(g["loss"]))
    # gradients against the float64 run of the reference class (its fp32 CPU run is itself only good to ~1e-4 here)
    assert _rel(x.grad.cpu().double(), torch.from_numpy(g["f64/dx"])) <= 2e-4
    for k, p in net.named_parameters():
        assert p.grad is not None, k
        assert _rel(p.grad.cpu().double(), torch.from_numpy(g["f64/grad/" + k])) <= 2e-4, k
    for k, v in net.state_dict().items():
        if "running" in k:
            torch.testing.assert_close(v.cpu(), torch.from_numpy(g["after/" + k]), rtol=1e-4, atol=1e-6, msg=k)
        if "num_batches" in k:
            assert int(v) == int(g["after/" + k])
    net.eval()
    with torch.no_grad():
        out_eval = net(x.detach())
    assert _rel(out_eval.cpu(), torch.from_numpy(g["out_eval"])) <= 1e-4


def test_discriminator_second_shape_against_oracle_and_bf16_tracks():
    """ndf 16, two strided blocks, non-cubic volume; oracle = plain-torch restatement on the module's own state_dict"""
    from synthanatomy_b200.networks.discriminator import B200Discriminator
    torch.manual_seed(5)
    net = B200Discriminator(input_nc=1, ndf=16, n_layers=2).cuda().train()
    state = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    for k, v in state.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    x = torch.rand(3, 1, 24, 40, 16)
    xr = x.clone().requires_grad_(True)
    ref = do.forward(state, xr, training=True)
    (ref ** 2).mean().backward()
    xd = x.cuda().requires_grad_(True)
    out = net(xd)
    (out ** 2).mean().backward()
    assert _rel(out.detach().cpu(), ref.detach()) <= 1e-4
    assert _rel(xd.grad.cpu(), xr.grad) <= 2e-4
    for k, p in net.named_parameters():
        assert _rel(p.grad.cpu(), state[k].grad) <= 2e-4, k
    # bf16 activations (what autocast selects): same function within bf16 rounding through four conv layers
    net16 = B200Discriminator(input_nc=1, ndf=16, n_layers=2, compute_dtype=torch.bfloat16).cuda().train()
    net16.load_state_dict({k: v.detach() for k, v in state.items()})
    out16 = net16(x.cuda())
    assert _rel(out16.detach().cpu(), ref.detach()) <= 6e-2
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert net._dtype() == torch.bfloat16


def test_discriminator_generator_side_gradient_only():
    """the generator's adversarial term: gradient w.r.t. the input with the discriminator's parameters frozen"""
    from synthanatomy_b200.networks.discriminator import B200Discriminator
    torch.manual_seed(6)
    net = B200Discriminator(input_nc=1, ndf=8, n_layers=3).cuda().train()
    state = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    for p in net.parameters():
        p.requires_grad_(False)
    x = torch.rand(2, 1, 32, 32, 32)
    xr = x.clone().requires_grad_(True)
    (-do.forward(state, xr, training=True).mean()).backward()
    xd = x.cuda().requires_grad_(True)
    (-net(xd).mean()).backward()
    assert _rel(xd.grad.cpu(), xr.grad) <= 2e-4
    assert all(p.grad is None for p in net.parameters())


@pytest.mark.parametrize("adaptive,amp", [(True, False), (False, True)])
def test_adversarial_iteration_with_the_dropin_networks(adaptive, amp):
    """one AdversarialTrainer iteration (trainer.py:122-262) over the drop-in VQ-VAE and discriminator: the adaptive weight
    differentiates the generator's graph three times (retain_activations), autocast selects the bf16 kernels"""
    from synthanatomy_b200 import engines
    from synthanatomy_b200.losses import MSELoss, get_discriminator_loss, get_generator_loss
    from synthanatomy_b200.networks.discriminator import B200Discriminator
    from synthanatomy_b200.networks.vqvae import B200VQVAE
    torch.manual_seed(9)
    G = B200VQVAE(n_levels=1, downsample_parameters=((4, 2, 1, 1),), upsample_parameters=((4, 2, 1, 0, 1),), n_embed=64,
                  embed_dim=8, n_channels=32, n_res_channels=32, n_res_layers=1, vq_decay=0.5, commitment_cost=0.25).cuda()
    D = B200Discriminator(input_nc=1, ndf=8, n_layers=2).cuda()
    og, od = torch.optim.Adam(G.parameters(), lr=1e-3), torch.optim.Adam(D.parameters(), lr=1e-3)
    g0 = [p.detach().clone() for p in G.parameters()]
    d0 = [p.detach().clone() for p in D.parameters()]
    x = torch.rand(2, 1, 16, 16, 16, device="cuda")
    out = engines.adversarial_iteration(x, x, G, D, og, od, MSELoss(), get_generator_loss({"generator_loss": "least_square"}),
                                        get_discriminator_loss({"discriminator_loss": "least_square"}), epoch=1,
                                        use_adversarial_adaptive_weight=adaptive, amp=amp)
    assert all(np.isfinite(out[k]) for k in ("loss", "g_loss", "d_loss"))
    assert out["g_loss"] >= out["loss"] - 1e-6                      # reconstruction + w * (non-negative adversarial term)
    assert any(not torch.equal(a, b.detach()) for a, b in zip(g0, G.parameters()))
    assert any(not torch.equal(a, b.detach()) for a, b in zip(d0, D.parameters()))
    assert all(torch.isfinite(p).all() for p in list(G.parameters()) + list(D.parameters()))
    if not adaptive:
        # without retain_activations a second differentiation of the generator's graph is refused, not silently wrong
        pred = G(x)
        loss = pred["reconstruction"][0].mean()
        torch.autograd.grad(loss, G.get_last_layer(), retain_graph=True)
        with pytest.raises(RuntimeError, match="retain_activations"):
            loss.backward()
