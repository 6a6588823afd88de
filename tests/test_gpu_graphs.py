"""CUDA-graph replay of the training step (synthanatomy_b200.utils.graphs.GraphedTrainStep): the replayed step is the eager
step -- same kernels, same order -- so the parameter trajectory of a graphed run must equal the eager one (up to the
fp32 atomics' ordering noise)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _close_params(a, b, tol):
    for (k, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        err = float((p.detach() - q.detach()).abs().max())
        assert err <= tol * max(1.0, float(p.detach().abs().max())), f"{k}: {err:.3e}"


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
    # eager twin: 2 warm-up steps on xs[0] (what the constructor below runs), then one step per batch
    opt_e = Adam(nets[0].parameters(), lr=1e-3)
    eager_losses = []
    for x in [xs[0], xs[0]] + xs:
        opt_e.zero_grad(set_to_none=True)
        loss = crit(nets[0](x), x)
        loss.backward()
        opt_e.step()
        eager_losses.append(float(loss))
    opt_g = Adam(nets[1].parameters(), lr=1e-3)
    step = GraphedTrainStep(nets[1], crit, opt_g, (xs[0],), xs[0], warmup=2)
    graph_losses = [float(step(x, target=x)) for x in xs]
    for a, b in zip(eager_losses[2:], graph_losses):
        assert abs(a - b) <= 1e-3 * abs(a), (eager_losses, graph_losses)
    _close_params(nets[0], nets[1], 2e-3)
    torch.testing.assert_close(nets[0].quantizer[0].impl.weight, nets[1].quantizer[0].impl.weight, rtol=1e-3, atol=1e-4)


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
    opt_e = Adam(nets[0].parameters(), lr=1e-3)
    eager_losses = []
    for x, y in [(toks[0], tgts[0])] * 2 + list(zip(toks, tgts)):
        opt_e.zero_grad(set_to_none=True)
        loss = crit(fwd(nets[0], x), y)
        loss.backward()
        opt_e.step()
        eager_losses.append(float(loss))
    opt_g = Adam(nets[1].parameters(), lr=1e-3)
    step = GraphedTrainStep(nets[1], crit, opt_g, (toks[0],), tgts[0], warmup=2, forward=fwd)
    graph_losses = [float(step(x, target=y)) for x, y in zip(toks, tgts)]
    for a, b in zip(eager_losses[2:], graph_losses):
        assert abs(a - b) <= 2e-3 * abs(a), (eager_losses, graph_losses)
    _close_params(nets[0], nets[1], 5e-3)
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
