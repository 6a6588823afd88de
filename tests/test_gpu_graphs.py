"""CUDA-graph replay of the training step (synthanatomy_b200.utils.graphs.GraphedTrainStep): the replayed step is the eager
step -- same kernels, same order.  Checked with the learning rate at 0 (so that Adam's sign-like first steps cannot amplify
the fp32 atomics' ordering noise into different parameters): per batch, the replayed loss and every parameter gradient
equal the eager ones; then, with the learning rate switched on (the optimiser runs outside the graph and reads it per
step), the loss falls under replay."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _close_grads(a, b, tol):
    for (k, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        if p.grad is None and q.grad is None:
            continue
        err = float((p.grad - q.grad).abs().max())
        assert err <= tol * max(float(p.grad.abs().max()), 1e-6), f"{k}: {err:.3e} of {float(p.grad.abs().max()):.3e}"


def test_graphed_vqvae_step_equals_eager():
    from synthanatomy_b200.losses import MSELoss
    from synthanatomy_b200.networks.vqvae import B200VQVAE
    from synthanatomy_b200.optim import Adam
    from synthanatomy_b200.utils.graphs import GraphedTrainStep
    kw = dict(n_levels=1, downsample_parameters=((4, 2, 1, 1),), upsample_parameters=((4, 2, 1, 0, 1),),
              n_embed=64, embed_dim=16, n_channels=128, n_res_channels=128, n_res_layers=2, vq_decay=0.5,
              commitment_cost=0.25, compute_dtype=torch.bfloat16)
    nets = []
    for _ in range(2):
        torch.manual_seed(0)
        nets.append(B200VQVAE(**kw).cuda().train())
    g = torch.Generator(device="cuda").manual_seed(1)
    xs = [torch.rand(2, 1, 16, 16, 32, device="cuda", generator=g) for _ in range(3)]
    crit = MSELoss()
    opt_e = Adam(nets[0].parameters(), lr=0.0)
    opt_g = Adam(nets[1].parameters(), lr=0.0)
    step = GraphedTrainStep(nets[1], crit, opt_g, (xs[0],), xs[0], warmup=2)
    for _ in range(2):                                   # keep the EMA codebooks in step with the constructor's warm-up
        crit(nets[0](xs[0]), xs[0]).backward()
    for x in xs:
        opt_e.zero_grad(set_to_none=True)
        loss = crit(nets[0](x), x)
        loss.backward()
        gl = step(x, target=x)
        assert abs(float(loss) - float(gl)) <= 1e-4 * abs(float(loss)), (float(loss), float(gl))
        _close_grads(nets[0], nets[1], 2e-3)
    torch.testing.assert_close(nets[0].quantizer[0].impl.weight, nets[1].quantizer[0].impl.weight, rtol=1e-3, atol=1e-4)
    opt_g.param_groups[0]["lr"] = 1e-3
    losses = [float(step(xs[0], target=xs[0])) for _ in range(8)]
    assert losses[-1] < losses[0], losses


def test_graphed_performer_step_equals_eager_and_redraws_outside_the_graph():
    from synthanatomy_b200.losses import CELoss
    from synthanatomy_b200.networks.transformers import Ordering, Performer
    from synthanatomy_b200.optim import Adam
    from synthanatomy_b200.utils.graphs import GraphedTrainStep
    grid = (4, 5, 6)
    n = int(np.prod(grid))
    order = Ordering("raster_scan", 3, (1, *grid), (False,) * 3, ((2, 0, 1),), ((0, 1),), ("rotate_90", "transpose"))
    kw = dict(num_tokens=65, dim=128, depth=2, heads=4, dim_head=64, local_attn_heads=2, local_window_size=20,
              max_seq_len=n + 1, ordering=order, causal=True, feature_redraw_interval=1, use_rezero=True,
              spatial_position_emb="absolute", spatial_shape=grid, compute_dtype=torch.bfloat16)
    nets = []
    for _ in range(2):
        torch.manual_seed(0)
        net = Performer(**kw).cuda().train()
        net.fix_projection_matrices_()
        nets.append(net)
    g = torch.Generator(device="cuda").manual_seed(2)
    toks = [torch.randint(0, 64, (2, n), device="cuda", generator=g) for _ in range(3)]
    tgts = [torch.randint(0, 64, (2, n), device="cuda", generator=g) for _ in range(3)]
    crit = CELoss()
    fwd = lambda m, x: m(x).transpose(1, 2)          # TransformerTrainingInferer
    opt_e = Adam(nets[0].parameters(), lr=0.0)
    opt_g = Adam(nets[1].parameters(), lr=0.0)
    step = GraphedTrainStep(nets[1], crit, opt_g, (toks[0],), tgts[0], warmup=2, forward=fwd)
    for x, y in zip(toks, tgts):
        opt_e.zero_grad(set_to_none=True)
        loss = crit(fwd(nets[0], x), y)
        loss.backward()
        gl = step(x, target=y)
        assert abs(float(loss) - float(gl)) <= 1e-4 * abs(float(loss)), (float(loss), float(gl))
        _close_grads(nets[0], nets[1], 2e-3)
    opt_g.param_groups[0]["lr"] = 1e-3
    losses = [float(step(toks[0], target=tgts[0])) for _ in range(8)]
    assert losses[-1] < losses[0], losses
    step.release()
    # with redraws: the projection matrices change between replays (copied into the static buffers outside the graph)
    torch.manual_seed(3)
    net = Performer(**kw).cuda().train()
    opt = Adam(net.parameters(), lr=1e-3)
    step = GraphedTrainStep(net, crit, opt, (toks[0],), tgts[0], warmup=2, forward=fwd, before_step=net.check_redraw_projections)
    proj = net.performer.net.layers[0][0].fn.fast_attention.projection_matrix
    seen = []
    for x, y in zip(toks, tgts):
        loss = step(x, target=y)
        assert torch.isfinite(loss)
        seen.append(proj.detach().clone())
    assert any(not torch.equal(seen[0], s) for s in seen[1:]), "no projection redraw happened between replays"


def test_graphed_step_in_deterministic_mode_replays_bit_identically():
    """the deterministic mode's ordered-sum buffers (stream-ordered cudaMallocAsync / cudaFreeAsync partials, zeroed counter
    slots) are captured with the step: two replays on the same batch at lr = 0 give bit-identical losses and gradients, and
    they equal a second, independently captured graph (Performer with fixed projections: no state changes under replay)"""
    from synthanatomy_b200 import ops
    from synthanatomy_b200.losses import CELoss
    from synthanatomy_b200.networks.transformers import Ordering, Performer
    from synthanatomy_b200.optim import Adam
    from synthanatomy_b200.utils.graphs import GraphedTrainStep
    grid = (6, 7, 8)
    n = int(np.prod(grid))
    order = Ordering("raster_scan", 3, (1, *grid), (False,) * 3, ((2, 0, 1),), ((0, 1),), ("rotate_90", "transpose"))
    kw = dict(num_tokens=65, dim=128, depth=2, heads=4, dim_head=64, local_attn_heads=2, local_window_size=20,
              max_seq_len=n + 1, ordering=order, causal=True, feature_redraw_interval=1, use_rezero=True,
              spatial_position_emb="absolute", spatial_shape=grid, compute_dtype=torch.bfloat16)
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randint(0, 64, (3, n), device="cuda", generator=g)
    y = torch.randint(0, 64, (3, n), device="cuda", generator=g)
    fwd = lambda m, t: m(t).transpose(1, 2)          # TransformerTrainingInferer
    ops.set_deterministic(True)
    try:
        results = []
        for _ in range(2):
            torch.manual_seed(0)
            net = Performer(**kw).cuda().train()
            net.fix_projection_matrices_()
            with torch.no_grad():
                for layer in net.performer.net.layers:
                    layer[0].g.fill_(0.5); layer[1].g.fill_(0.5)
            opt = Adam(net.parameters(), lr=0.0)
            step = GraphedTrainStep(net, CELoss(), opt, (x,), y, warmup=2, forward=fwd)
            runs = []
            for _ in range(2):
                loss = step(x, target=y).detach().clone()
                runs.append((loss, [p.grad.clone() for p in net.parameters() if p.grad is not None]))
            assert torch.isfinite(runs[0][0])
            assert torch.equal(runs[0][0], runs[1][0])
            for a, b in zip(runs[0][1], runs[1][1]):
                assert torch.equal(a, b)
            results.append(runs[0])
            step.release()
        assert torch.equal(results[0][0], results[1][0])
        for a, b in zip(results[0][1], results[1][1]):
            assert torch.equal(a, b)
    finally:
        ops.set_deterministic(None)
