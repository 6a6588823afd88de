"""Generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

The only shim is a ``sys.modules`` stub for ``monai.networks.blocks.SubpixelUpsample``
(imported at baseline.py:6 but only used when ``use_subpixel_conv=True``, baseline.py:274-282).
TEST INFRASTRUCTURE -- never imported by the product.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference():
    for name in ("monai", "monai.networks", "monai.networks.blocks"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["monai.networks.blocks"].SubpixelUpsample = type("SubpixelUpsample", (), {})
    sys.path.insert(0, REF)
    from src.networks.vqvae.baseline import BaselineVQVAE, Quantizer_impl  # noqa
    return BaselineVQVAE, Quantizer_impl


def np_sd(sd):
    return {k: v.detach().cpu().numpy().copy() for k, v in sd.items()}


def vqvae_case(BaselineVQVAE, name, seed, n_levels, ch, vol, batch, n_embed=2048, embed_dim=32, codebook_scale=None):
    torch.manual_seed(seed)
    net = BaselineVQVAE(
        n_levels=n_levels,
        downsample_parameters=((4, 2, 1, 1),) * n_levels,
        upsample_parameters=((4, 2, 1, 0, 1),) * n_levels,
        n_embed=n_embed, embed_dim=embed_dim, n_channels=ch, n_res_channels=ch, n_res_layers=3,
        vq_decay=0.5, commitment_cost=0.25,
    )
    # the default nn.Embedding init N(0,1) leaves nearly every code unused for 32-dim relu-free latents of
    # magnitude ~0.1; keep the reference init (it is what the reference does) -- ties are covered by vq cases.
    if codebook_scale is not None:  # spread the latents over many codes (a loadable state, not a code change)
        with torch.no_grad():
            net.quantizer[0].impl.embedding.weight.mul_(codebook_scale)
            net.quantizer[0].impl.embed_avg.copy_(net.quantizer[0].impl.embedding.weight)
    x = torch.rand(batch, 1, *vol)
    sd0 = np_sd(net.state_dict())
    net.train()
    out = net(x)
    recon = out["reconstruction"][0]
    q_loss = out["quantization_losses"][0]
    loss = F.mse_loss(recon.float(), x.float()) + q_loss.float()
    loss.backward()
    grads = {k: p.grad.detach().numpy().copy() for k, p in net.named_parameters() if p.grad is not None}
    sd1 = np_sd(net.state_dict())
    # eval-mode API slices (run_vqvae.py extracting / decoding modes) with the POST-update codebook
    net.eval()
    with torch.no_grad():
        idx = net.index_quantize(x)[0]
        dec = net.decode_samples([idx])
        enc = net.encode(x)[0]
    blob = {"x": x.numpy(), "recon": recon.detach().numpy(), "q_loss": q_loss.detach().numpy(),
            "loss": loss.detach().numpy(), "perplexity": net.get_perplexity()[0].detach().numpy(),
            "eval_idx": idx.numpy(), "eval_decode": dec.numpy(), "eval_encode": enc.numpy(),
            "cfg": np.array([n_levels, ch, n_embed, embed_dim, batch, *vol], dtype=np.int64)}
    for k, v in sd0.items():
        if k in ("quantizer.0.impl.embedding.weight", "quantizer.0.impl.embed_avg"):
            assert np.array_equal(v, sd0["quantizer.0.impl.weight"])  # aliases at init (baseline.py:33,36)
            continue
        blob["sd0/" + k] = v
    for k, v in grads.items():
        blob["grad/" + k] = v
    for k in ("quantizer.0.impl.weight", "quantizer.0.impl.N", "quantizer.0.impl.embed_avg"):
        blob["sd1/" + k] = sd1[k]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)
    print(name, "loss", float(loss), "perplexity", float(blob["perplexity"]),
          "unique idx", len(np.unique(idx.numpy())), "params", sum(v.size for v in grads.values()))


def vq_cases(Quantizer_impl):
    """BASELINE config 3 + adversarial sets.  Inputs that are derived from (z, W) by a
    recipe are stored as the recipe's integer draws, and rebuilt by tests/golden_util.py the same way."""
    K, D = 2048, 32
    g0 = torch.Generator().manual_seed(0)
    g1 = torch.Generator().manual_seed(1)
    z = torch.randn(8, 32, 10, 14, 10, generator=g0)
    W = torch.randn(K, D, generator=g1)
    blob = {"z": z.numpy(), "W": W.numpy()}

    def run(tag, z, W, steps=1, training=True):
        q = Quantizer_impl(K, D, 1e-5)
        with torch.no_grad():
            q.embedding.weight.copy_(W)
            q.embed_avg.copy_(W)
        q.train(training)
        for s in range(steps):
            qst, loss, idx = q(z, 0.5, 0.25)
            assert int(idx.max()) < 32768
            blob[f"{tag}/idx{s}"] = idx.numpy().astype(np.int16)
            blob[f"{tag}/loss{s}"] = loss.numpy()
        if training:
            blob[f"{tag}/N"] = q.N.numpy().copy()
            blob[f"{tag}/embed_avg"] = q.embed_avg.numpy().copy()
            blob[f"{tag}/weight"] = q.weight.detach().numpy().copy()

    run("plain", z, W, steps=3)                       # 3 EMA steps (codebook moves between steps)
    W2 = W.clone(); W2[7] = W2[3]; W2[100] = W2[3]; W2[2047] = W2[0]
    run("dup", z, W2, steps=1, training=False)         # duplicate rows -> lowest index wins
    rows = 2 * 10 * 14 * 10
    sel = torch.randint(0, K, (rows,), generator=g0)
    exact = W[sel].reshape(2, 10, 14, 10, D).permute(0, 4, 1, 2, 3).contiguous()
    blob["exact/sel"] = sel.numpy().astype(np.int16)
    run("exact", exact, W, steps=1, training=False)   # latents exactly equal to codebook rows
    i = torch.randint(0, K, (rows,), generator=g0)
    j = torch.randint(0, K, (rows,), generator=g0)
    noise = 1e-7 * torch.randn(rows, D, generator=g0)
    near = ((W[i] + W[j]) / 2 + noise).reshape(2, 10, 14, 10, D).permute(0, 4, 1, 2, 3).contiguous()
    blob["near/i"] = i.numpy().astype(np.int16); blob["near/j"] = j.numpy().astype(np.int16)
    blob["near/noise"] = noise.numpy()
    run("near", near, W, steps=1, training=False)     # near-ties between two codes
    small = torch.randn(1, 32, 2, 3, 1, generator=g0)  # ragged tiny grid
    blob["tiny/z"] = small.numpy()
    run("tiny", small, W, steps=2)
    np.savez_compressed(os.path.join(OUT, "vq_cfg3.npz"), **blob)
    print("vq_cfg3 written;", {k: v.shape for k, v in blob.items() if k.endswith("idx0")})


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # deterministic reduction order for the fixtures
    BaselineVQVAE, Quantizer_impl = import_reference()
    # BASELINE.json configs[0]: 1-level, 32 ch, 32^3, batch 2
    vqvae_case(BaselineVQVAE, "vqvae_cfg1", 0, 1, 32, (32, 32, 32), 2)
    # multi-level indexing / channel halving / non-cubic volume; small codebook so codes are reused
    vqvae_case(BaselineVQVAE, "vqvae_l2", 4, 2, 16, (16, 24, 8), 1, n_embed=64, embed_dim=8, codebook_scale=0.05)
    vq_cases(Quantizer_impl)


if __name__ == "__main__":
    main()
