"""CPU: pin oracle/vqvae_oracle.py against the fixtures produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from oracle import vqvae_oracle as vo
from tests import golden_util as gu


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)   # the goldens were made single-threaded (fixed reduction order)
    yield
    torch.set_num_threads(n)


@pytest.mark.parametrize("name", ["vqvae_cfg1", "vqvae_l2"])
def test_forward_backward_matches_reference(name):
    kw, sd, blob = gu.vqvae_case(name)
    cfg = vo.VQVAEConfig(**kw)
    x = torch.from_numpy(blob["x"])
    loss, grads, out = vo.train_step_grads(sd, cfg, x)
    np.testing.assert_array_equal(out["reconstruction"][0].detach().numpy(), blob["recon"])
    np.testing.assert_array_equal(out["quantization_losses"][0].detach().numpy(), blob["q_loss"])
    np.testing.assert_array_equal(loss.numpy(), blob["loss"])
    gold = {k[5:]: v for k, v in blob.items() if k.startswith("grad/")}
    assert set(gold) == set(grads)
    for k, g in gold.items():
        np.testing.assert_allclose(grads[k].numpy(), g, rtol=0, atol=1e-7, err_msg=k)
    ns = out["new_state"]
    np.testing.assert_array_equal(ns["N"].numpy(), blob["sd1/quantizer.0.impl.N"])
    np.testing.assert_array_equal(ns["embed_avg"].numpy(), blob["sd1/quantizer.0.impl.embed_avg"])
    np.testing.assert_array_equal(ns["weight"].numpy(), blob["sd1/quantizer.0.impl.weight"])
    np.testing.assert_array_equal(vo.perplexity(out["indices"], cfg.n_embed).numpy(), blob["perplexity"])


@pytest.mark.parametrize("name", ["vqvae_cfg1", "vqvae_l2"])
def test_eval_api_slices(name):
    kw, sd, blob = gu.vqvae_case(name)
    cfg = vo.VQVAEConfig(**kw)
    sd = dict(sd)
    sd["quantizer.0.impl.weight"] = torch.from_numpy(blob["sd1/quantizer.0.impl.weight"])
    x = torch.from_numpy(blob["x"])
    with torch.no_grad():
        z = vo.encode(sd, cfg, x)
        _, _, idx, ns = vo.quantize(sd, cfg, z, training=False)
        dec = vo.decode(sd, cfg, vo.embed(sd, idx))
    assert ns is None
    np.testing.assert_array_equal(z.numpy(), blob["eval_encode"])
    np.testing.assert_array_equal(idx.numpy(), blob["eval_idx"])
    np.testing.assert_array_equal(dec.numpy(), blob["eval_decode"])


@pytest.mark.parametrize("tag,steps,training", [("plain", 3, True), ("dup", 1, False), ("exact", 1, False),
                                                 ("near", 1, False), ("tiny", 2, True)])
def test_quantizer_cases(tag, steps, training):
    blob = gu.load("vq_cfg3")
    z, W = gu.vq_inputs(blob, tag)
    cfg = vo.VQVAEConfig(n_embed=W.shape[0], embed_dim=W.shape[1])
    sd = {"quantizer.0.impl.weight": W.clone(), "quantizer.0.impl.N": torch.zeros(W.shape[0]),
          "quantizer.0.impl.embed_avg": W.clone()}
    for s in range(steps):
        _, loss, idx, ns = vo.quantize(sd, cfg, z, training=training)
        np.testing.assert_array_equal(idx.numpy(), blob[f"{tag}/idx{s}"].astype(np.int64))
        np.testing.assert_array_equal(loss.numpy(), blob[f"{tag}/loss{s}"])
        if training:
            sd = {"quantizer.0.impl." + k: v for k, v in ns.items()}
    if training:
        for k in ("N", "embed_avg", "weight"):
            np.testing.assert_array_equal(sd["quantizer.0.impl." + k].numpy(), blob[f"{tag}/{k}"])
    if tag == "dup":  # duplicated rows 3==7==100 and 0==2047: the lowest index must win
        got = blob["dup/idx0"]
        assert not np.isin(got, [7, 100, 2047]).any()


def test_naive_conv_pins_functional():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 3, 5, 6, 4, generator=g)
    for k, s, p in [(3, 1, 1), (4, 2, 1), (1, 1, 0)]:
        w = torch.randn(4, 3, k, k, k, generator=g)
        b = torch.randn(4, generator=g)
        ref = torch.nn.functional.conv3d(x, w, b, stride=s, padding=p).numpy()
        np.testing.assert_allclose(vo.conv3d_naive(x.numpy(), w.numpy(), b.numpy(), s, p), ref, atol=1e-4)
    wt = torch.randn(3, 2, 4, 4, 4, generator=g)
    bt = torch.randn(2, generator=g)
    ref = torch.nn.functional.conv_transpose3d(x, wt, bt, stride=2, padding=1).numpy()
    np.testing.assert_allclose(vo.conv_transpose3d_naive(x.numpy(), wt.numpy(), bt.numpy(), 2, 1), ref, atol=1e-4)


def test_program_indices_match_survey():
    cfg = vo.VQVAEConfig(n_levels=4, downsample_parameters=((4, 2, 1, 1),) * 4,
                         upsample_parameters=((4, 2, 1, 0, 1),) * 4, n_channels=256, n_res_channels=256,
                         n_embed=2048, embed_dim=32)
    assert [i for _, i, _ in vo.encoder_program(cfg)] == [0, 2, 3, 5, 6, 8, 9, 11, 12]
    assert [i for _, i, _ in vo.decoder_program(cfg)] == [0, 1, 2, 4, 5, 7, 8, 10, 11]
    sd = vo.init_state_dict(cfg)
    trainable = sum(v.numel() for k, v in sd.items() if not k.startswith("quantizer."))
    assert trainable == 28123937
