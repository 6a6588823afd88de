"""The vendor-library comparator of bench.py (tools/vendor_baseline.py: stock torch modules -> cuDNN / cuBLAS on the
GPU box) computes the same functions as the pinned CPU oracles -- checked here on small shapes, on the CPU."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import performer_oracle as po
from oracle import vqvae_oracle as vo
from tools import vendor_baseline as vb


def test_vendor_vqvae_equals_oracle_step():
    kw = dict(n_levels=2, n_embed=64, embed_dim=8, n_channels=16, n_res_layers=2)
    cfg = vo.VQVAEConfig(n_levels=2, downsample_parameters=((4, 2, 1, 1),) * 2, upsample_parameters=((4, 2, 1, 0, 1),) * 2,
                         n_embed=64, embed_dim=8, n_channels=16, n_res_channels=16, n_res_layers=2, vq_decay=0.5)
    sd = vo.init_state_dict(cfg, 1)
    sd["quantizer.0.impl.weight"] = sd["quantizer.0.impl.weight"] * 0.05
    sd["quantizer.0.impl.embedding.weight"] = sd["quantizer.0.impl.weight"]
    sd["quantizer.0.impl.embed_avg"] = sd["quantizer.0.impl.weight"].clone()
    net = vb.VendorVQVAE(**kw).train()
    convs = [m for m in net.modules() if isinstance(m, (nn.Conv3d, nn.ConvTranspose3d))]
    keys = [k[:-7] for k in sd if k.endswith(".weight") and not k.startswith("quantizer.")]
    assert len(convs) == len(keys)
    with torch.no_grad():
        for m, k in zip(convs, keys):
            m.weight.copy_(sd[k + ".weight"]); m.bias.copy_(sd[k + ".bias"])
        net.weight.copy_(sd["quantizer.0.impl.weight"]); net.embed_avg.copy_(sd["quantizer.0.impl.embed_avg"])
    x = torch.rand(2, 1, 16, 16, 16, generator=torch.Generator().manual_seed(2))
    loss_ref, grads_ref, out_ref = vo.train_step_grads(sd, cfg, x)
    rec, ql = net(x)
    loss = F.mse_loss(rec, x) + ql
    loss.backward()
    torch.testing.assert_close(rec, out_ref["reconstruction"][0], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(loss, loss_ref, rtol=1e-5, atol=1e-7)
    for m, k in zip(convs, keys):
        torch.testing.assert_close(m.weight.grad, grads_ref[k + ".weight"], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(net.weight, out_ref["new_state"]["weight"], rtol=1e-5, atol=1e-6)


def test_vendor_performer_equals_oracle_forward():
    grid = (3, 4, 5)
    n = int(np.prod(grid))
    kw = dict(num_tokens=33, dim=64, depth=2, heads=4, dim_head=64, local_attn_heads=2, local_window_size=16)
    cfg = po.PerformerConfig(max_seq_len=n + 1, spatial_shape=grid, **kw)
    sd = po.init_state_dict(cfg, 3)
    for i in range(cfg.depth):
        sd[po.layer_prefix(i) + "0.g"] = torch.tensor(0.7); sd[po.layer_prefix(i) + "1.g"] = torch.tensor(-0.4)
    order = np.arange(n)                                                     # plain raster order
    seqs = [torch.from_numpy(s.copy()) for s in po.spatial_index_sequences(grid, order)]
    tok = torch.randint(0, 32, (2, n), generator=torch.Generator().manual_seed(4))
    want = po.forward(sd, cfg, tok, seqs)
    P = {"tok": sd["token_emb.weight"], "pos": sd["pos_emb.emb.weight"],
         "sp": [sd[f"spatial_position_emb.{a}.emb.weight"] for a in range(3)], "nw": sd["norm.weight"], "nb": sd["norm.bias"],
         "Wout": sd["to_out.weight"], "bout": sd["to_out.bias"], "layers": []}
    for i in range(cfg.depth):
        p = po.layer_prefix(i)
        P["layers"].append({"ga": sd[p + "0.g"], "gf": sd[p + "1.g"], "Wq": sd[p + "0.fn.to_q.weight"],
                            "Wk": sd[p + "0.fn.to_k.weight"], "Wv": sd[p + "0.fn.to_v.weight"], "Wo": sd[p + "0.fn.to_out.weight"],
                            "W1": sd[p + "1.fn.fn.w1.weight"], "b1": sd[p + "1.fn.fn.w1.bias"], "W2": sd[p + "1.fn.fn.w2.weight"],
                            "b2": sd[p + "1.fn.fn.w2.bias"], "proj": sd[p + "0.fn.fast_attention.projection_matrix"]})
    sp_idx = torch.stack([torch.cat((torch.full((1,), -1, dtype=torch.long), s[: n - 1].long())) for s in seqs])
    got = vb.vendor_performer_forward(P, tok, sp_idx, 4, 2, 64, 16)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)
