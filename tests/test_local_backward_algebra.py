"""The algebra the local-attention backward kernels implement (csrc/sa_tc_local.cu, sa_pf_local_simt.cu), restated tile by
tile in float64 and checked against autograd through the oracle's bucketed local attention (oracle/performer_oracle.py,
following local-attention's LocalAttention.forward as built by performer-pytorch 1.0.11 for
/root/reference/src/networks/transformers/performer.py:194-219):

    window of query p:   (floor(p / w) - 1) * w <= j <= p
    forward, flash style over key tiles:  running max / sum,  lse_p = max + log sum,  O = sum_j P_pj v_j
    delta_p = dO_p . O_p                                        (handed from the dq kernel to the dk/dv kernel)
    dS_pj = P_pj (dO_p . v_j - delta_p),   P_pj = exp(d^-1/2 q_p . k_j - lse_p)   (recomputed from lse, never stored)
    dq_p = d^-1/2 sum_j dS_pj k_j,   dk_j = d^-1/2 sum_p dS_pj q_p,   dv_j = sum_p P_pj dO_p
    rotary term: q, k are rotated in place before the forward; dq, dk leave through the TRANSPOSE of the rotation
        y1 = x1 c - x2 s, y2 = x2 c + x1 s   =>   g1' = g1 c + g2 s, g2' = g2 c - g1 s        (pairs (e, e + d/2))

No GPU: this pins the derivation (including which key tiles a query tile visits), the GPU tests pin the kernels."""
import math

import pytest
import torch

from oracle import performer_oracle as po


def _rot(x, freqs):            # the forward rotation on [B, H, N, d] (cos / sin of the oracle's fp32 angle table, as it forms them)
    return x * freqs.cos() + po.rotate_half(x) * freqs.sin()


def _rot_t(g, freqs):          # its transpose, written on the (e, e + d/2) pairs as the kernels' epilogues do
    h = g.shape[-1] // 2
    c, s = freqs.cos()[..., :h], freqs.sin()[..., :h]
    g1, g2 = g[..., :h], g[..., h:]
    return torch.cat((g1 * c + g2 * s, g2 * c - g1 * s), dim=-1)


def _tiled(q, k, v, dout, W, tile):
    """forward + backward over (query tile, key tile) pairs, visiting only the tiles a window can touch"""
    B, H, N, d = q.shape
    sc = d ** -0.5
    out, lse = torch.zeros_like(q), q.new_zeros(B, H, N)
    pos = torch.arange(N)
    lo = (pos // W - 1).clamp(min=0) * W
    allowed = (pos[None, :] <= pos[:, None]) & (pos[None, :] >= lo[:, None])
    visited = 0
    for p0 in range(0, N, tile):
        ps = slice(p0, min(p0 + tile, N))
        mx = q.new_full((B, H, ps.stop - p0), -math.inf)
        sm = q.new_zeros(B, H, ps.stop - p0)
        acc = q.new_zeros(B, H, ps.stop - p0, d)
        first = (int(lo[p0]) // tile) * tile                      # first key tile the first query of the tile can see
        for j0 in range(first, ps.stop, tile):
            js = slice(j0, min(j0 + tile, N))
            visited += 1
            s = torch.einsum("bhpe,bhje->bhpj", q[:, :, ps], k[:, :, js]) * sc
            s = s.masked_fill(~allowed[ps, js], -math.inf)
            new = torch.maximum(mx, s.amax(-1))
            safe = torch.where(torch.isinf(new), torch.zeros_like(new), new)
            corr = torch.exp(mx - safe)
            pexp = torch.exp(s - safe.unsqueeze(-1))
            sm = sm * corr + pexp.sum(-1)
            acc = acc * corr.unsqueeze(-1) + torch.einsum("bhpj,bhje->bhpe", pexp, v[:, :, js])
            mx = new
        out[:, :, ps] = acc / sm.unsqueeze(-1)
        lse[:, :, ps] = mx + sm.log()
    delta = (dout * out).sum(-1)
    dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
    for p0 in range(0, N, tile):
        ps = slice(p0, min(p0 + tile, N))
        first = (int(lo[p0]) // tile) * tile
        for j0 in range(first, ps.stop, tile):
            js = slice(j0, min(j0 + tile, N))
            s = torch.einsum("bhpe,bhje->bhpj", q[:, :, ps], k[:, :, js]) * sc
            P = torch.exp(s - lse[:, :, ps].unsqueeze(-1)).masked_fill(~allowed[ps, js], 0.0)
            dS = P * (torch.einsum("bhpe,bhje->bhpj", dout[:, :, ps], v[:, :, js]) - delta[:, :, ps].unsqueeze(-1))
            dq[:, :, ps] += sc * torch.einsum("bhpj,bhje->bhpe", dS, k[:, :, js])
            dk[:, :, js] += sc * torch.einsum("bhpj,bhpe->bhje", dS, q[:, :, ps])
            dv[:, :, js] += torch.einsum("bhpj,bhpe->bhje", P, dout[:, :, ps])
    return out, dq, dk, dv, visited


@pytest.mark.parametrize("N,W,tile,rel", [(50, 12, 8, "rotary"), (37, 7, 16, "rotary"), (64, 16, 16, "none"), (19, 20, 8, "rotary"),
                                          (45, 5, 16, "none")])
def test_tiled_forward_and_backward_equal_autograd(N, W, tile, rel):
    g = torch.Generator().manual_seed(N + W)
    B, H, d = 1, 2, 8
    q, k, v = (torch.randn(B, H, N, d, generator=g, dtype=torch.float64).requires_grad_(True) for _ in range(3))
    dout = torch.randn(B, H, N, d, generator=g, dtype=torch.float64)
    out_ref = po.local_attention(q, k, v, W, rel)
    (out_ref * dout).sum().backward()
    with torch.no_grad():
        qd, kd = q.detach(), k.detach()
        if rel == "rotary":
            freqs = po.sinusoidal_embeddings(N, d)[None, None]
            qd, kd = _rot(qd, freqs), _rot(kd, freqs)
        out, dq, dk, dv, visited = _tiled(qd, kd, v.detach(), dout, W, tile)
        if rel == "rotary":
            dq, dk = _rot_t(dq, freqs), _rot_t(dk, freqs)
    for name, got, want in (("out", out, out_ref.detach()), ("dq", dq, q.grad), ("dk", dk, k.grad), ("dv", dv, v.grad)):
        err = float((got - want).abs().max())
        assert err <= 1e-10 * max(1.0, float(want.abs().max())), f"{name}: {err:.3e}"
    ntiles = -(-N // tile)
    assert visited <= ntiles * (2 + -(-2 * W // tile)), "a query tile visits only the key tiles of its own and the previous window"
