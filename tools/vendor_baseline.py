"""Vendor-library comparator for bench.py (BASELINE.md section 4 last row): the reference's two
models written with stock torch modules / functions, so that on a B200 every hot op lands in cuDNN 9 / cuBLAS / ATen --
the kernels the reference's own Python would reach on this box with ``cudnn.benchmark = True``
(/root/reference/src/utils/general.py:336-338).  NOT part of the product and never imported by it: bench.py times it
next to the product so that the speed-up over the vendor sm_100 kernels is visible in the same JSON line.

  * ``VendorVQVAE``: the layer list of /root/reference/src/networks/vqvae/baseline.py:150-160, 213-299 as
    nn.Conv3d / nn.ConvTranspose3d / nn.ReLU, the quantiser of :38-87 in torch ops (distance matrix, one-hot, EMA);
  * ``vendor_performer_*``: the Performer layer of SURVEY.md section 10 in torch ops (F.linear, einsum, softmax); the
    causal numerator, which the reference takes from the fast-transformers CUDA extension (absent here), as the chunked
    einsum form.

tests/test_vendor_baseline.py pins both to the CPU oracles on small shapes, so the comparator computes the same function.
"""
from __future__ import annotations

import math
import time
from typing import Dict, List, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ VQ-VAE
class _Res(nn.Module):
    def __init__(self, c, cr):
        super().__init__()
        self.res = nn.Sequential(nn.Conv3d(c, cr, 3, padding=1), nn.ReLU(True), nn.Dropout3d(0.0), nn.Conv3d(cr, c, 1))

    def forward(self, x):
        return F.relu(x + self.res(x))


class VendorVQVAE(nn.Module):
    def __init__(self, n_levels=4, n_embed=2048, embed_dim=32, n_channels=256, n_res_layers=3, commitment_cost=0.25,
                 vq_decay=0.5, eps=1e-5):
        super().__init__()
        enc: List[nn.Module] = []
        for i in range(n_levels):
            last = i == n_levels - 1
            c = n_channels // (1 if last else 2)
            enc += [nn.Conv3d(1 if i == 0 else n_channels // 2, c, 4, 2, 1), nn.ReLU(),
                    nn.Sequential(*[_Res(c, c) for _ in range(n_res_layers)])]
        enc.append(nn.Conv3d(n_channels, embed_dim, 3, 1, 1))
        self.encoder = nn.Sequential(*enc)
        dec: List[nn.Module] = [nn.Conv3d(embed_dim, n_channels, 3, 1, 1)]
        for i in range(n_levels):
            first, last = i == 0, i == n_levels - 1
            c = n_channels // (1 if first else 2)
            dec.append(nn.Sequential(*[_Res(c, c) for _ in range(n_res_layers)]))
            dec.append(nn.ConvTranspose3d(c, 1 if last else n_channels // 2, 4, 2, 1))
            if not last:
                dec.append(nn.ReLU())
        self.decoder = nn.Sequential(*dec)
        self.register_buffer("weight", torch.randn(n_embed, embed_dim))
        self.register_buffer("N", torch.zeros(n_embed))
        self.register_buffer("embed_avg", self.weight.clone())
        self.n_embed, self.beta, self.decay, self.eps = n_embed, commitment_cost, vq_decay, eps

    def quantize(self, x):
        with torch.autocast(x.device.type, enabled=False):            # baseline.py:38
            b, c, h, w, d = x.shape
            x = x.float()
            flat = x.permute(0, 2, 3, 4, 1).contiguous().view(-1, c)
            dist = (flat ** 2).sum(1, keepdim=True) - 2 * flat @ self.weight.t() + (self.weight ** 2).sum(1)[None]
            idx = (-dist).max(1)[1]
            onehot = F.one_hot(idx, self.n_embed).type_as(flat)
            q = F.embedding(idx.view(b, h, w, d), self.weight).permute(0, 4, 1, 2, 3).contiguous()
            if self.training:
                with torch.no_grad():
                    self.N.mul_(self.decay).add_(onehot.sum(0), alpha=1 - self.decay)
                    self.embed_avg.mul_(self.decay).add_(onehot.t() @ flat, alpha=1 - self.decay)
                    n = self.N.sum()
                    W = (self.N + self.eps) / (n + self.n_embed * self.eps) * n
                    self.weight.copy_(self.embed_avg / W[:, None])
            loss = self.beta * F.mse_loss(q.detach(), x)
            return (q - x).detach() + x, loss

    def forward(self, x):
        q, loss = self.quantize(self.encoder(x))
        return self.decoder(q), loss


def time_vendor_vqvae(batch: int, vol: Sequence[int], steps: int, warmup: int, dtype, kw: Dict) -> Dict:
    """fwd + bwd + Adam of the stock-torch model on cuda; returns ms/step (CUDA events) and the batch that fitted"""
    dev = torch.device("cuda")
    torch.backends.cudnn.benchmark = True                  # src/utils/general.py:336-338
    torch.backends.cudnn.allow_tf32 = True                 # NGC default of the reference's image (TF32 convs)
    torch.backends.cuda.matmul.allow_tf32 = True
    b = batch
    while b >= 1:
        net = opt = x = None
        try:
            torch.manual_seed(4)
            net = VendorVQVAE(**kw).to(dev).train().to(memory_format=torch.channels_last_3d)
            opt = torch.optim.Adam(net.parameters(), lr=1.65e-4)
            x = torch.rand(b, 1, *vol, device=dev)

            def step():
                with torch.autocast("cuda", dtype=dtype, enabled=dtype is not None):
                    rec, ql = net(x)
                    loss = F.mse_loss(rec.float(), x) + ql
                loss.backward()
                opt.step()
                opt.zero_grad(set_to_none=True)
                return loss

            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            return {"ms_per_step": ms, "batch": b, "value": b / (ms / 1e3), "loss": float(loss)}
        except torch.OutOfMemoryError:
            b //= 2
        finally:
            del net, opt, x
            torch.cuda.empty_cache()
    return {"unavailable": "out of memory at batch 1"}


# ------------------------------------------------------------------------------------------------ Performer
def _softmax_kernel(data, proj, is_query, eps=1e-4):
    d = data.shape[-1]
    c = d ** -0.25
    ratio = proj.shape[0] ** -0.5
    dash = torch.einsum("...id,jd->...ij", c * data, proj)
    diag = ((data ** 2).sum(-1) / 2.0 * (c ** 2)).unsqueeze(-1)
    if is_query:
        return ratio * (torch.exp(dash - diag - dash.max(dim=-1, keepdim=True).values) + eps)
    return ratio * (torch.exp(dash - diag - dash.max()) + eps)


def _causal_linear_attention(q, k, v, eps=1e-6, chunk=128):
    k_cumsum = k.cumsum(dim=-2) + eps
    d_inv = 1.0 / torch.einsum("...nd,...nd->...n", q, k_cumsum)
    B, H, N, m = q.shape
    state = q.new_zeros(B, H, m, v.shape[-1])
    outs = []
    for s in range(0, N, chunk):
        qc, kc, vc = q[:, :, s:s + chunk], k[:, :, s:s + chunk], v[:, :, s:s + chunk]
        a = torch.einsum("bhim,bhjm->bhij", qc, kc).tril()
        outs.append(torch.einsum("bhij,bhje->bhie", a, vc) + torch.einsum("bhim,bhme->bhie", qc, state))
        state = state + torch.einsum("bhjm,bhje->bhme", kc, vc)
    return torch.cat(outs, dim=2) * d_inv.unsqueeze(-1)


def _rotary(q, k):
    n, d = q.shape[-2], q.shape[-1]
    inv_freq = 1.0 / (10000 ** (torch.arange(0, d, 2, device=q.device).float() / d))
    f = torch.einsum("i,j->ij", torch.arange(n, device=q.device).float(), inv_freq)
    f = torch.cat((f, f), dim=-1)[None]

    def rot(x):
        x1, x2 = x[..., : d // 2], x[..., d // 2:]
        return torch.cat((-x2, x1), dim=-1)

    return q * f.cos() + rot(q) * f.sin(), k * f.cos() + rot(k) * f.sin()


def _look_around(x, pad_value):
    t = x.shape[1]
    dims = (0, 0) * (x.dim() - 2)
    padded = F.pad(x, (*dims, 1, 0), value=pad_value)
    return torch.cat((padded[:, 0:t], padded[:, 1:t + 1]), dim=2)


def _local_attention(q, k, v, w):
    shape = q.shape
    q, k, v = (t.reshape(-1, *t.shape[-2:]) for t in (q, k, v))
    q, k = _rotary(q, k)
    n0 = q.shape[1]
    rem = (-n0) % w
    if rem:
        q, k, v = (F.pad(t, (0, 0, 0, rem), value=0.0) for t in (q, k, v))
    b, t, e = q.shape
    nw = t // w
    tick = torch.arange(t, device=q.device, dtype=q.dtype).reshape(1, nw, w)
    bq, bk, bv = (x.reshape(b, nw, w, -1) for x in (q, k, v))
    bk, bv = _look_around(bk, -1.0), _look_around(bv, -1.0)
    tk = _look_around(tick, -1.0)
    dots = torch.einsum("bhie,bhje->bhij", bq, bk) * (e ** -0.5)
    neg = -torch.finfo(dots.dtype).max
    dots = dots.masked_fill(tick[:, :, :, None] < tk[:, :, None, :], neg)
    dots = dots.masked_fill(tk[:, :, None, :] == -1, neg)
    out = torch.einsum("bhij,bhje->bhie", dots.softmax(dim=-1), bv).reshape(-1, t, e)
    return out[:, :n0].reshape(*shape)


def vendor_performer_init(dim, depth, heads, dim_head, num_tokens, n, grid, device, seed=4):
    g = torch.Generator().manual_seed(seed)
    inner = heads * dim_head

    def lin(o, i):
        b = 1.0 / math.sqrt(i)
        return ((torch.rand(o, i, generator=g) * 2 - 1) * b).to(device).requires_grad_(True)

    m = int(dim_head * math.log(dim_head))
    P = {"tok": torch.randn(num_tokens, dim, generator=g).to(device).requires_grad_(True),
         "pos": torch.randn(n + 1, dim, generator=g).to(device).requires_grad_(True),
         # one table per axis with N rows, indexed by coordinate value (performer.py:27-33: nn.Embedding(len(seq) - 1, dim))
         "sp": [torch.randn(n, dim, generator=g).to(device).requires_grad_(True) for _ in grid],
         "layers": [], "nw": torch.ones(dim, device=device, requires_grad=True),
         "nb": torch.zeros(dim, device=device, requires_grad=True), "Wout": lin(num_tokens, dim),
         "bout": torch.zeros(num_tokens, device=device, requires_grad=True)}
    for _ in range(depth):
        P["layers"].append({
            "ga": torch.tensor(1e-3, device=device, requires_grad=True), "gf": torch.tensor(1e-3, device=device, requires_grad=True),
            "Wq": lin(inner, dim), "Wk": lin(inner, dim), "Wv": lin(inner, dim), "Wo": lin(dim, inner),
            "W1": lin(4 * dim, dim), "b1": torch.zeros(4 * dim, device=device, requires_grad=True),
            "W2": lin(dim, 4 * dim), "b2": torch.zeros(dim, device=device, requires_grad=True),
            "proj": torch.randn(m, dim_head, generator=g).to(device)})
    return P


def vendor_performer_params(P):
    out = [P["tok"], P["pos"], *P["sp"], P["nw"], P["nb"], P["Wout"], P["bout"]]
    for L in P["layers"]:
        out += [L[k] for k in ("ga", "gf", "Wq", "Wk", "Wv", "Wo", "W1", "b1", "W2", "b2")]
    return out


def vendor_performer_forward(P, tokens, sp_idx, heads, local_heads, dim_head, window):
    """tokens [B, N] int64; sp_idx [n_axes, N] (coordinate of position n - 1, -1 at BOS) -> logits [B, N, V]"""
    B, N = tokens.shape
    x = F.embedding(tokens, P["tok"]) + P["pos"][:N]
    for a, tab in enumerate(P["sp"]):
        row = sp_idx[a]
        x = x + torch.where((row >= 0)[:, None], tab[row.clamp(min=0)], torch.zeros((), device=x.device))
    gh = heads - local_heads
    for L in P["layers"]:
        q, k, v = (F.linear(x, L[n]).reshape(B, N, heads, dim_head).permute(0, 2, 1, 3) for n in ("Wq", "Wk", "Wv"))
        outs = []
        if gh > 0:
            with torch.autocast(x.device.type, enabled=False):        # exp / cumsum / 1/x in fp32, as the reference
                qp = _softmax_kernel(q[:, :gh].float(), L["proj"], True)
                kp = _softmax_kernel(k[:, :gh].float(), L["proj"], False)
                outs.append(_causal_linear_attention(qp, kp, v[:, :gh].float()).to(q.dtype))
        if local_heads > 0:
            outs.append(_local_attention(q[:, gh:], k[:, gh:], v[:, gh:], window))
        att = torch.cat(outs, dim=1).permute(0, 2, 1, 3).reshape(B, N, heads * dim_head)
        x = x + F.linear(att, L["Wo"]) * L["ga"]
        x = x + F.linear(F.gelu(F.linear(x, L["W1"], L["b1"])), L["W2"], L["b2"]) * L["gf"]
    x = F.layer_norm(x, (x.shape[-1],), P["nw"], P["nb"], 1e-5)
    return F.linear(x, P["Wout"], P["bout"])


def time_vendor_performer(batch, grid, depth, steps, warmup, dtype, kw) -> Dict:
    """fwd + CE + bwd + Adam on cuda through cuBLAS / ATen; batch halves on OOM (autograd keeps the materialised feature
    and score tensors of every layer, which the reference's fast-transformers extension would not)"""
    import numpy as np
    dev = torch.device("cuda")
    torch.backends.cuda.matmul.allow_tf32 = True           # the reference's effective mode: fp32 storage, TF32 matmuls
    torch.backends.cudnn.allow_tf32 = True
    n = int(np.prod(grid))
    heads, lh, dh, w, V, dim = kw["heads"], kw["local_attn_heads"], kw["dim_head"], kw["local_window_size"], kw["num_tokens"], kw["dim"]
    coords = np.stack(np.meshgrid(*[np.arange(s) for s in grid], indexing="ij")).reshape(len(grid), -1)
    sp_idx = torch.from_numpy(np.concatenate([np.full((len(grid), 1), -1), coords[:, : n - 1]], axis=1)).to(dev)
    b = batch
    while b >= 1:
        P = opt = None
        try:
            P = vendor_performer_init(dim, depth, heads, dh, V, n, grid, dev)
            opt = torch.optim.Adam(vendor_performer_params(P), lr=1e-3)
            tok = torch.randint(0, V - 1, (b, n), device=dev)
            tgt = torch.randint(0, V - 1, (b, n), device=dev)

            def step():
                with torch.autocast("cuda", dtype=dtype, enabled=dtype is not None):
                    logits = vendor_performer_forward(P, tok, sp_idx, heads, lh, dh, w)
                loss = F.cross_entropy(logits.transpose(1, 2).float(), tgt)
                loss.backward()
                opt.step()
                opt.zero_grad(set_to_none=True)
                return loss

            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                loss = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            return {"ms_per_step": ms, "batch": b, "value": b * n / (ms / 1e3), "loss": float(loss)}
        except torch.OutOfMemoryError:
            b //= 2
        finally:
            del P, opt
            torch.cuda.empty_cache()
    return {"unavailable": "out of memory at batch 1"}


if __name__ == "__main__":
    t0 = time.time()
    print(time_vendor_vqvae(8, (160, 224, 160), 3, 2, torch.bfloat16,
                            dict(n_levels=4, n_embed=2048, embed_dim=32, n_channels=256, n_res_layers=3)))
    print(time_vendor_performer(6, (20, 28, 25), 24, 2, 1, None,
                                dict(heads=16, local_attn_heads=8, dim_head=64, local_window_size=420, num_tokens=2049,
                                     dim=512)))
    print(f"{time.time() - t0:.1f} s")
