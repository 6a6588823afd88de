"""CPU: the Performer oracle (oracle/performer_oracle.py) and the product's host-side logic.

Pinned against the unmodified reference (tests/golden/performer_host.npz, made by oracle/make_golden_performer.py):
the ordering index sequences and prepare_batch.  The third-party arithmetic is "parity unpinned" (see the oracle's
header); what is checked here is that its independent formulations agree with each other and have the documented
properties of the reference's algorithm."""
import numpy as np
import pytest
import torch

from oracle import make_golden_performer as mg
from oracle import performer_oracle as po
from tests import golden_util as gu


@pytest.fixture(scope="module")
def host():
    return gu.load("performer_host")


@pytest.mark.parametrize("case", mg.ORDER_CASES, ids=[c[0] for c in mg.ORDER_CASES])
def test_ordering_matches_reference(host, case):
    from synthanatomy_b200.networks.transformers import Ordering
    name, typ, dims, refl, tr, rot, order = case
    want = host[f"order/{name}"]
    got_oracle = po.ordering_restated(typ, dims, refl, tr, rot, order)
    o = Ordering(typ, len(dims) - 1, dims, refl, tr, rot, order)
    assert np.array_equal(got_oracle, want)
    assert np.array_equal(o.get_sequence_ordering(), want)
    assert np.array_equal(o.get_revert_sequence_ordering(), host[f"revert/{name}"])
    x = torch.arange(len(want))
    assert torch.equal(o(x), torch.from_numpy(want.astype(np.int64)))


def test_prepare_batch_matches_reference(host):
    from synthanatomy_b200.utils.transformer import prepare_batch
    q = host["pb/quantization"]
    seq = host["order/readme_10x14x10"].astype(np.int64)
    xi, y = po.prepare_batch(q, seq, 2048)
    assert np.array_equal(xi, host["pb/x_input"].astype(np.int64))
    assert np.array_equal(y, host["pb/x_target"].astype(np.int64))
    (xi2, cond), y2 = prepare_batch({"quantization": torch.from_numpy(q)}, seq, 2048)
    assert cond is None and xi2.dtype == torch.int64
    assert np.array_equal(xi2.numpy(), xi) and np.array_equal(y2.numpy(), y)
    assert (xi[:, 0] == 2048).all()          # BOS = vocab_size on the left


def test_causal_prefix_sum_equals_masked_quadratic_form():
    g = torch.Generator().manual_seed(0)
    q, k = torch.rand(2, 3, 150, 20, generator=g), torch.rand(2, 3, 150, 20, generator=g)
    v = torch.randn(2, 3, 150, 8, generator=g)
    a = torch.einsum("bhim,bhjm->bhij", q, k).tril()
    want = a @ v
    for chunk in (1, 7, 64, 150, 1000):
        torch.testing.assert_close(po.causal_dot_product(q, k, v, chunk=chunk), want, rtol=1e-5, atol=1e-4)   # |values| ~ 100
    out = po.causal_linear_attention(q, k, v)
    den = a.sum(-1, keepdim=True) + 1e-6 * q.sum(-1, keepdim=True)
    torch.testing.assert_close(out, want / den, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("n,w", [(50, 20), (40, 20), (19, 20), (101, 7)])
@pytest.mark.parametrize("rel", ["rotary", "none"])
def test_local_attention_bucketed_equals_dense(n, w, rel):
    g = torch.Generator().manual_seed(1)
    q, k, v = (torch.randn(2, 3, n, 16, generator=g) for _ in range(3))
    torch.testing.assert_close(po.local_attention(q, k, v, w, rel), po.local_attention_dense(q, k, v, w, rel),
                               rtol=1e-5, atol=1e-6)


def test_local_heads_are_strictly_causal_and_global_heads_couple_through_the_key_max():
    cfg = po.PerformerConfig(num_tokens=33, max_seq_len=61, dim=32, depth=1, heads=4, dim_head=16, local_attn_heads=4,
                             local_window_size=10, spatial_shape=None, spatial_position_emb=None)
    sd = po.init_state_dict(cfg, 3)
    sd["performer.net.layers.0.0.g"] = torch.tensor(1.0)
    tok = torch.randint(0, 32, (1, 60), generator=torch.Generator().manual_seed(5))
    tok2 = tok.clone(); tok2[0, -1] = (tok2[0, -1] + 1) % 32
    a, b = po.forward(sd, cfg, tok), po.forward(sd, cfg, tok2)
    assert torch.equal(a[:, :-1], b[:, :-1])            # local heads only: changing the last token changes nothing earlier
    # global heads: the stabiliser is one max over the whole key tensor (performer-pytorch 1.0.11)
    g = torch.Generator().manual_seed(7)
    k = torch.randn(1, 2, 30, 16, generator=g)
    P = po.gaussian_orthogonal_random_matrix(40, 16, generator=g)
    k2 = k.clone(); k2[0, 0, -1] *= 4.0                  # moves the global max
    f1, f2 = po.softmax_kernel(k, P, False), po.softmax_kernel(k2, P, False)
    assert not torch.allclose(f1[0, 1], f2[0, 1])       # another head's features changed
    f3 = po.softmax_kernel(k2, P, False, key_stabiliser="per_head")
    f4 = po.softmax_kernel(k, P, False, key_stabiliser="per_head")
    assert torch.equal(f3[0, 1], f4[0, 1])


def test_projection_matrix_blocks_are_orthogonal_and_redraw_rule():
    g = torch.Generator().manual_seed(0)
    P = po.gaussian_orthogonal_random_matrix(266, 64, generator=g)
    assert P.shape == (266, 64)
    blk = P[:64] / P[:64].norm(dim=1, keepdim=True)
    torch.testing.assert_close(blk @ blk.t(), torch.eye(64), rtol=0, atol=1e-5)
    upd = po.ProjectionUpdaterState(1)
    assert [upd.step(True) for _ in range(6)] == [False, True, False, True, False, True]   # every 2nd training forward
    assert upd.step(False) is False
    assert po.PerformerConfig().m == 266


def test_product_module_tree_matches_oracle_keys_and_redraw_rule():
    from synthanatomy_b200.networks.transformers import Ordering, Performer
    o = Ordering("raster_scan", 3, (1, 4, 5, 6), (False,) * 3, ((2, 0, 1),), ((0, 1),), ("rotate_90", "transpose"))
    net = Performer(num_tokens=65, max_seq_len=121, dim=64, depth=2, heads=4, ordering=o, dim_head=64, local_attn_heads=2,
                    local_window_size=20, feature_redraw_interval=1, use_rezero=True, spatial_position_emb="absolute",
                    spatial_shape=(4, 5, 6))
    cfg = po.PerformerConfig(num_tokens=65, max_seq_len=121, dim=64, depth=2, heads=4, dim_head=64, local_attn_heads=2,
                             local_window_size=20, spatial_shape=(4, 5, 6))
    sd = po.init_state_dict(cfg, 0)
    mine = net.state_dict()
    for k, v in sd.items():
        assert k in mine and tuple(mine[k].shape) == tuple(v.shape), k
    extra = set(mine) - set(sd)
    assert all(("proj_updater" in k) or k.endswith(("inv_freq", "spatial_indices_sequence")) for k in extra), extra
    # the duplicated keys through ProjectionUpdater.instance exist in performer-pytorch 1.0.11 as well
    assert "performer.proj_updater.instance.layers.0.0.fn.to_q.weight" in mine
    n_train = sum(p.numel() for p in net.parameters())
    assert n_train == sum(sd[k].numel() for k in po.trainable_keys(sd))
    net.train()
    P0 = net.performer.net.layers[0][0].fn.fast_attention.projection_matrix.clone()
    net.performer.proj_updater.redraw_projections()      # call 1: counts
    assert torch.equal(P0, net.performer.net.layers[0][0].fn.fast_attention.projection_matrix)
    net.performer.proj_updater.redraw_projections()      # call 2: redraws
    assert not torch.equal(P0, net.performer.net.layers[0][0].fn.fast_attention.projection_matrix)
    net.fix_projection_matrices_()
    P1 = net.performer.net.layers[0][0].fn.fast_attention.projection_matrix.clone()
    for _ in range(4):
        net.check_redraw_projections()
    assert torch.equal(P1, net.performer.net.layers[0][0].fn.fast_attention.projection_matrix)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 8, dtype=torch.long))
    with pytest.raises(NotImplementedError):
        Performer(num_tokens=65, max_seq_len=121, dim=64, depth=1, heads=4, ordering=o, use_rezero=False)


def test_readme_parameter_count():
    """105.7 M parameters at N = 1400 (24 layers x 4 196 866 + embeddings + head)."""
    cfg = po.PerformerConfig(spatial_shape=(10, 14, 10))
    per_layer = 4 * 1024 * 512 + (2048 * 512 + 2048) + (512 * 2048 + 512) + 2
    assert per_layer == 4196866
    total = 24 * per_layer + 2049 * 512 + 1401 * 512 + 3 * 1399 * 512 + 2 * 512 + 2049 * 512 + 2049
    assert abs(total - 105.7e6) < 0.1e6


def test_restatement_against_installed_packages():
    """the third-party arithmetic of the Performer oracle (performer-pytorch 1.0.11, local-attention)
    is restated from the published source, not executed -- neither package is in this image, in /opt/wheelhouse or on the
    GPU box, so this test SKIPS there.  Wherever the packages can be imported it pins the restatement to them:
    softmax_kernel (query / key forms), the non-CUDA causal linear attention, the local window attention."""
    import pytest
    pp = pytest.importorskip("performer_pytorch.performer_pytorch")
    g = torch.Generator().manual_seed(0)
    data = torch.randn(2, 3, 37, 64, generator=g)
    proj = torch.randn(40, 64, generator=g)
    for is_query in (True, False):
        want = pp.softmax_kernel(data, projection_matrix=proj, is_query=is_query)
        torch.testing.assert_close(po.softmax_kernel(data, proj, is_query), want, rtol=1e-5, atol=1e-7)
    q = po.softmax_kernel(data, proj, True)
    k = po.softmax_kernel(data, proj, False)
    v = torch.randn(2, 3, 37, 64, generator=g)
    if hasattr(pp, "causal_linear_attention_noncuda"):
        want = pp.causal_linear_attention_noncuda(q, k, v)
        torch.testing.assert_close(po.causal_linear_attention(q, k, v), want, rtol=1e-4, atol=1e-6)
    la = pytest.importorskip("local_attention")
    attn = la.LocalAttention(window_size=8, causal=True, autopad=True, look_backward=1, look_forward=0, dropout=0.0,
                             rel_pos_emb_config=(64, 3))
    x = [torch.randn(2, 3, 37, 64, generator=g) for _ in range(3)]
    want = attn(*x)
    rel = "rotary" if any("rel_pos" in n or "inv_freq" in n for n, _ in attn.named_buffers()) else "none"
    torch.testing.assert_close(po.local_attention(*x, 8, rel), want, rtol=1e-4, atol=1e-6)
