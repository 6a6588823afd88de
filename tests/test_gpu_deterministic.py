"""Deterministic mode (the reference's `deterministic=True`, run_vqvae.py:550 / run_transformer.py:417 ->
src/utils/general.py:333 sets torch.backends.cudnn.deterministic): with the switch on, every cross-CTA floating-point sum of
the library is added in one fixed order, so two runs of the same training step give bit-identical losses and gradients.
The default mode adds split-K partials in arrival order (atomics) and is only reproducible to rounding."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture
def deterministic():
    from synthanatomy_b200 import ops
    ops.set_deterministic(True)
    yield
    ops.set_deterministic(None)


def _vq_grads(dt, shape, seed=0):
    from synthanatomy_b200.losses import MSELoss
    from synthanatomy_b200.networks.vqvae import B200VQVAE
    torch.manual_seed(seed)
    net = B200VQVAE(n_levels=2, downsample_parameters=((4, 2, 1, 1),) * 2, upsample_parameters=((4, 2, 1, 0, 1),) * 2,
                    n_embed=64, embed_dim=32, n_channels=128, n_res_channels=128, n_res_layers=2, vq_decay=0.5,
                    commitment_cost=0.25, compute_dtype=dt).cuda().train()
    x = torch.rand(*shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(seed + 1))
    out = net(x)
    loss = MSELoss()(out, x)
    loss.backward()
    grads = [p.grad.clone() for p in net.parameters() if p.grad is not None]
    assert len(grads) > 20
    extra = [b.clone() for b in net.buffers()]
    return loss.detach().clone(), grads, extra


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
def test_vqvae_training_step_is_bit_reproducible(dt, deterministic):
    """2-level 128-channel VQ-VAE, 2 x 1 x 64 x 64 x 64 (bf16: tcgen05 conv / split-K weight-gradient / fused pointwise
    kernels with many position splits; fp32: the CUDA-core kernels): loss, every gradient and the EMA codebook buffers"""
    from synthanatomy_b200.ops import lib
    a = _vq_grads(dt, (2, 1, 64, 64, 64))
    assert lib().sa_get_deterministic() == 1
    b = _vq_grads(dt, (2, 1, 64, 64, 64))
    assert torch.equal(a[0], b[0])
    for i, (ga, gb) in enumerate(zip(a[1], b[1])):
        assert torch.equal(ga, gb), f"gradient {i} differs by {float((ga - gb).abs().max())}"
    for i, (ba, bb) in enumerate(zip(a[2], b[2])):
        assert torch.equal(ba, bb), f"buffer {i} differs"


def test_default_mode_matches_deterministic_mode_to_rounding():
    """the ordered sums are the same sums: the two modes agree to fp32 accumulation-order noise"""
    from synthanatomy_b200 import ops
    ops.set_deterministic(False)
    try:
        a = _vq_grads(torch.bfloat16, (2, 1, 64, 64, 64))
        ops.set_deterministic(True)
        b = _vq_grads(torch.bfloat16, (2, 1, 64, 64, 64))
    finally:
        ops.set_deterministic(None)
    assert abs(float(a[0]) - float(b[0])) <= 1e-5 * abs(float(b[0]))
    for ga, gb in zip(a[1], b[1]):
        assert float((ga - gb).abs().max()) <= 1e-4 * max(float(gb.abs().max()), 1e-6)


def test_switch_follows_torch_flag():
    from synthanatomy_b200 import ops
    from synthanatomy_b200.ops import lib
    ops.set_deterministic(None)
    old = torch.backends.cudnn.deterministic
    try:
        torch.backends.cudnn.deterministic = True
        assert ops.sync_deterministic() is True and lib().sa_get_deterministic() == 1
        torch.backends.cudnn.deterministic = False
        assert ops.sync_deterministic() is False and lib().sa_get_deterministic() == 0
    finally:
        torch.backends.cudnn.deterministic = old
        ops.sync_deterministic()


def _pf_grads(dt, seed=0):
    import numpy as np
    from synthanatomy_b200.losses import CELoss
    from synthanatomy_b200.networks.transformers import Ordering, Performer
    from synthanatomy_b200.utils.transformer import prepare_batch
    grid = (10, 14, 10)
    n = int(np.prod(grid))
    torch.manual_seed(seed)
    order = Ordering("raster_scan", 3, (1, *grid), (False,) * 3, ((2, 0, 1),), ((0, 1),), ("rotate_90", "transpose"))
    net = Performer(num_tokens=2049, dim=512, heads=16, dim_head=64, local_attn_heads=8, local_window_size=420,
                    max_seq_len=n + 1, depth=2, ordering=order, causal=True, feature_redraw_interval=1000,
                    generalized_attention=False, use_rezero=True, spatial_position_emb="absolute", spatial_shape=grid,
                    compute_dtype=dt).cuda().train()
    with torch.no_grad():                      # open the ReZero gates so that every gradient is a real sum
        for layer in net.performer.net.layers:
            layer[0].g.fill_(0.5); layer[1].g.fill_(0.5)
    quant = torch.randint(0, 2048, (3, *grid), generator=torch.Generator().manual_seed(seed + 2))
    (x, _), y = prepare_batch({"quantization": quant}, order.get_sequence_ordering(), 2048)
    logits = net(x.cuda())
    loss = CELoss()(logits.transpose(1, 2), y.cuda())
    loss.backward()
    named = [(k, p.grad.clone()) for k, p in net.named_parameters() if p.grad is not None]
    assert len(named) > 20
    return loss.detach().clone(), named


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
def test_performer_training_step_is_bit_reproducible(dt, deterministic):
    """dim 512, 2 layers, 16 heads (8 local), 1400 tokens, batch 3: split-K weight-gradient GEMMs with their bias column
    sums, gate gradients, the FAVOR+ stabiliser gradient, LayerNorm / embedding-table gradients and the loss sum"""
    a = _pf_grads(dt)
    b = _pf_grads(dt)
    assert torch.equal(a[0], b[0])
    for (k, ga), (_, gb) in zip(a[1], b[1]):
        assert torch.equal(ga, gb), f"gradient {k} differs by {float((ga - gb).abs().max())}"
