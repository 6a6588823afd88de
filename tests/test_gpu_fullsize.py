"""Full-size checks (BASELINE.json configs 2 and 4) through size-independent properties: the CPU oracle cannot run these
shapes in seconds, so the CUDA path is checked against identities that hold at any size --

  * adjointness of the conv kernels at the level-1 shape  <conv(x), g> = <x, dgrad(g)> = <w, wgrad(x, g)>,
  * fused pointwise backward == the three general kernels it replaces,
  * tcgen05 (bf16) path against the CUDA-core path on the same tensors (FAVOR+ scan at 14 000 tokens),
  * determinism, encode -> index -> decode round trip and batch-permutation equivariance of the two full models.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from synthanatomy_b200 import ops
    return ops


def _dot(a, b):
    return float((a.double() * b.double()).sum())


def test_level1_conv_adjoint_identities_full_size():
    """3x3x3 128->128 at 8 x 80 x 112 x 80 (the dominant launch of config 2): forward, data gradient and weight
    gradient of the tcgen05 kernels are adjoint to one another (bf16 outputs: 3e-3 relative on ~1e8-term sums)."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(0)
    B, D, H, W, C = 8, 80, 112, 80, 128
    x = torch.randn(B, D, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
    gy = torch.randn(B, D, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(C, C, 3, 3, 3, device="cuda", generator=g) * 0.02).to(torch.bfloat16).float()
    spec = ops.ConvSpec("conv", C, C, 3, 1, 1)
    y = ops.conv_forward(spec, x, ops.pack_weight(w, False, torch.bfloat16), None, None, False)
    assert ops.last_path() == 2
    gy = (y.float() + 0.5 * gy.float()).to(torch.bfloat16)      # correlated with y: <y, g> ~ |y|^2, not a random-sign sum
    dx = ops.conv_dgrad(spec, gy, ops.pack_weight(w, True, torch.bfloat16), (D, H, W))
    dw = ops.conv_wgrad(spec, x, gy, w)
    lhs = _dot(y, gy)
    assert abs(lhs - _dot(x, dx)) <= 3e-3 * abs(lhs), (lhs, _dot(x, dx))
    assert abs(lhs - _dot(w, dw)) <= 3e-3 * abs(lhs), (lhs, _dot(w, dw))
    # linearity in the input (no bias, no ReLU): conv(2 x) == 2 conv(x) exactly in bf16 (power-of-two scaling)
    y2 = ops.conv_forward(spec, (x.float() * 2).to(torch.bfloat16), ops.pack_weight(w, False, torch.bfloat16), None, None, False)
    assert torch.equal(y2.float(), y.float() * 2)


def test_strided_pair_adjoint_full_size():
    """4/2/1 strided conv 128->128, 80 x 112 x 80 -> 40 x 56 x 40 (B = 8), and its transposed twin"""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(1)
    B, D, H, W, C = 8, 80, 112, 80, 128
    x = torch.randn(B, D, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
    gy = torch.randn(B, D // 2, H // 2, W // 2, C, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(C, C, 4, 4, 4, device="cuda", generator=g) * 0.02).to(torch.bfloat16).float()
    spec = ops.ConvSpec("conv", C, C, 4, 2, 1)
    y = ops.conv_forward(spec, x, ops.pack_weight(w, False, torch.bfloat16), None, None, False)
    gy = (y.float() + 0.5 * gy.float()).to(torch.bfloat16)
    dx = ops.conv_dgrad(spec, gy, ops.pack_weight(w, True, torch.bfloat16), (D, H, W))
    dw = ops.conv_wgrad(spec, x, gy, w)
    lhs = _dot(y, gy)
    assert abs(lhs - _dot(x, dx)) <= 3e-3 * abs(lhs), (lhs, _dot(x, dx))
    assert abs(lhs - _dot(w, dw)) <= 3e-3 * abs(lhs), (lhs, _dot(w, dw))


def test_pointwise_backward_fused_equals_general_kernels_full_size():
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(2)
    B, D, H, W, C = 8, 80, 112, 80, 128
    gy = torch.randn(B, D, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
    h = torch.relu(torch.randn(B, D, H, W, C, device="cuda", generator=g)).to(torch.bfloat16)
    w = (torch.randn(C, C, 1, 1, 1, device="cuda", generator=g) * 0.05).to(torch.bfloat16).float()
    spec = ops.ConvSpec("conv", C, C, 1, 1, 0)
    wp_t = ops.pack_weight(w, True, torch.bfloat16)
    dh, dw, db, dbh = ops.conv1x1_bwd_fused(spec, gy, h, wp_t, w, with_dbh=True)
    assert torch.equal(dh, ops.conv_dgrad(spec, gy, wp_t, (D, H, W), None, h))
    dbh2 = ops.bias_grad(dh)             # the separate streaming reduction the fused column sums replace
    torch.testing.assert_close(dbh, dbh2, rtol=1e-3, atol=1e-3 * float(dbh2.abs().max()))
    dw2, db2 = ops.conv_wgrad(spec, h, gy, w), ops.bias_grad(gy)
    torch.testing.assert_close(dw, dw2, rtol=1e-3, atol=1e-3 * float(dw2.abs().max()))
    torch.testing.assert_close(db, db2, rtol=1e-3, atol=1e-3 * float(db2.abs().max()))


def test_vqvae_config2_determinism_and_round_trip():
    """baseline_vqvae 4-level 256ch at 160 x 224 x 160 (B = 2): two evaluations agree bit for bit, and
    decode_samples(index_quantize(x)) reproduces forward's reconstruction (encode -> indices -> decode round trip)."""
    from synthanatomy_b200.networks.vqvae import B200VQVAE
    torch.manual_seed(4)
    net = B200VQVAE(n_levels=4, downsample_parameters=((4, 2, 1, 1),) * 4, upsample_parameters=((4, 2, 1, 0, 1),) * 4,
                    n_embed=2048, embed_dim=32, n_channels=256, n_res_channels=256, n_res_layers=3, vq_decay=0.5,
                    commitment_cost=0.25, compute_dtype=torch.bfloat16).cuda().eval()
    x = torch.rand(2, 1, 160, 224, 160, device="cuda")
    with torch.no_grad():
        out1 = net(x)["reconstruction"][0]
        out2 = net(x)["reconstruction"][0]
        idx = net.index_quantize(x)
        rec = net.decode_samples(idx)
    assert out1.shape == x.shape and torch.isfinite(out1).all()
    assert torch.equal(out1, out2)
    assert tuple(idx[0].shape) == (2, 10, 14, 10) and int(idx[0].min()) >= 0 and int(idx[0].max()) < 2048
    torch.testing.assert_close(rec, out1, rtol=0, atol=1e-5)
    # batch independence in eval mode: sample 0 alone gives the same reconstruction
    with torch.no_grad():
        solo = net(x[:1])["reconstruction"][0]
    assert torch.equal(solo, out1[:1])


def test_favor_scan_tc_matches_cuda_core_path_at_14000_tokens():
    """the tcgen05 causal scan (110 chunks, bf16 states) against the CUDA-core kernels on the same bf16 buffers"""
    from synthanatomy_b200 import ops, pf_ops as pf
    g = torch.Generator(device="cuda").manual_seed(3)
    B, H, N, m, mp, d = 1, 2, 14000, 266, 272, 64
    QF = torch.zeros(B, H, N, mp, device="cuda"); KF = torch.zeros(B, H, N, mp, device="cuda")
    QF[..., :m] = torch.rand(B, H, N, m, device="cuda", generator=g) * 0.1 + 1e-3
    KF[..., :m] = torch.rand(B, H, N, m, device="cuda", generator=g) * 0.1 + 1e-3
    QF, KF = QF.bfloat16(), KF.bfloat16()
    vbuf = torch.randn(B * N, H * d, device="cuda", generator=g).bfloat16()
    fd = pf.favor_desc(B, N, H, d, m, mp, H * d, torch.bfloat16)
    ws = torch.empty(pf.favor_scan_workspace(fd, True), dtype=torch.uint8, device="cuda")
    outs = []
    for simt in (False, True):
        ops.set_force_simt(simt)
        try:
            O = torch.zeros(B * N, H * d, device="cuda", dtype=torch.bfloat16)
            den = torch.empty(B, H, N, device="cuda")
            pf.favor_scan_fwd(fd, QF, KF, vbuf, 0, 1e-6, O, 0, den, ws)
            assert ops.last_path() == (1 if simt else 2)
            outs.append((O.float(), den))
        finally:
            ops.set_force_simt(False)
    torch.testing.assert_close(outs[0][1], outs[1][1], rtol=5e-3, atol=0)
    assert float((outs[0][0] - outs[1][0]).abs().max()) <= 2e-2 * float(outs[1][0].abs().max())


def test_performer_config4_batch_permutation_equivariance():
    """Performer dim 512, 24 layers, 16 heads (8 local, w 420), grid 20 x 28 x 25 = 14 000 tokens, batch 6 (eval mode):
    permuting the batch permutes the logits (the global key stabiliser is a batch-wide max, hence invariant)."""
    from synthanatomy_b200.networks.transformers import Ordering, Performer
    from synthanatomy_b200.utils.transformer import prepare_batch
    grid = (20, 28, 25)
    n = int(np.prod(grid))
    torch.manual_seed(4)
    order = Ordering("raster_scan", 3, (1, *grid), (False,) * 3, ((2, 0, 1),), ((0, 1),), ("rotate_90", "transpose"))
    net = Performer(num_tokens=2049, dim=512, heads=16, dim_head=64, local_attn_heads=8, local_window_size=420,
                    max_seq_len=n + 1, depth=24, ordering=order, causal=True, feature_redraw_interval=1,
                    generalized_attention=False, use_rezero=True, spatial_position_emb="absolute", spatial_shape=grid,
                    compute_dtype=torch.bfloat16).cuda().eval()
    with torch.no_grad():                      # open the ReZero gates: with g = 1e-3 the layers barely matter
        for layer in net.performer.net.layers:
            layer[0].g.fill_(0.5); layer[1].g.fill_(0.5)
    quant = torch.randint(0, 2048, (6, *grid), generator=torch.Generator().manual_seed(2))
    (x, _), _y = prepare_batch({"quantization": quant}, order.get_sequence_ordering(), 2048)
    x = x.cuda()
    perm = torch.tensor([3, 0, 5, 1, 4, 2], device="cuda")
    with torch.no_grad():
        a = net(x)
        b = net(x[perm])
    assert tuple(a.shape) == (6, n, 2049) and torch.isfinite(a).all()
    assert torch.equal(a[perm], b)


def test_local_attention_tc_matches_cuda_core_path_at_14000_tokens():
    """window 420, look-back one window, 14 000 tokens (110 query tiles, ~15 key tiles each): the tcgen05 kernels (O, P, dS
    resident in TMEM, lazily rescaled reference maximum) against the CUDA-core kernels on the same bf16 buffers"""
    from synthanatomy_b200 import ops, pf_ops as pf
    g = torch.Generator(device="cuda").manual_seed(7)
    B, H, N, d, W = 1, 2, 14000, 64, 420
    inner = H * d
    buf = (torch.randn(B * N, 3 * inner, device="cuda", generator=g) * 0.7).bfloat16()
    dout = torch.randn(B * N, inner, device="cuda", generator=g).bfloat16()
    desc = pf.local_desc(B, N, H, d, W, 3 * inner, inner, torch.bfloat16)
    res = []
    for simt in (False, True):
        ops.set_force_simt(simt)
        try:
            out = torch.zeros(B * N, inner, device="cuda", dtype=torch.bfloat16)
            lse = torch.empty(B * H, N, device="cuda")
            dbuf = torch.zeros_like(buf)
            pf.local_attn_fwd(desc, buf, 0, inner, 2 * inner, None, out, 0, lse)
            assert ops.last_path() == (1 if simt else 2)
            pf.local_attn_bwd(desc, buf, 0, inner, 2 * inner, None, out, dout, 0, lse, dbuf)
            res.append((out.float(), lse.clone(), dbuf.float()))
        finally:
            ops.set_force_simt(False)
    (o_tc, l_tc, d_tc), (o_ref, l_ref, d_ref) = res
    torch.testing.assert_close(l_tc, l_ref, rtol=0, atol=2e-3)
    assert float((o_tc - o_ref).abs().max()) <= 2e-2 * float(o_ref.abs().max())
    for i, name in enumerate(("dq", "dk", "dv")):
        a, b = d_tc[:, i * inner:(i + 1) * inner], d_ref[:, i * inner:(i + 1) * inner]
        assert float((a - b).abs().max()) <= 2e-2 * float(b.abs().max()), name
