"""The algebra the FAVOR+ backward kernels implement (csrc/sa_tc_favor.cu, sa_pf_favor_simt.cu),
restated chunk by chunk in float64 and checked against autograd through the oracle's softmax_kernel +
causal_linear_attention (oracle/performer_oracle.py, which follows performer-pytorch 1.0.11 / fast-transformers as called
from /root/reference/src/networks/transformers/performer.py:270):

    delta_i = dout_i . out_i,  inv_i = 1 / den_i
    dq'_i = inv_i (dout_i S^T - delta_i ksum + sum_{j <= i in chunk} (dout_i . v_j - delta_i) k'_j)           [tc_dqk<0>]
    dk'_j = v_j R^T + Rden + sum_{i >= j in chunk} ((v_j . dout_i) - delta_i) inv_i q'_i                       [tc_dqk<1>]
    dv_j  = k'_j R + sum_{i >= j in chunk} (k'_j . q'_i) inv_i dout_i                                          [tc_scan<1>]
    dD = df (feat - r eps),  s = row sum of dD                                                                 [featmap_bwd /
    dx = c (dD P - s P[argmax]) - c^2 s x       (queries: row stabiliser, not detached)                         tc_dqk_fb]
    dx = c dD P - c^2 s x;  dx[n*] -= c (sum of all s) P[f*]   (keys: one global stabiliser)                   [kmax_fixup]

S / ksum are the exclusive prefix states (sum_j k'_j (x) v_j, sum_j k'_j + eps) of the chunk, R / Rden the exclusive
suffix states (sum_i q'_i inv_i (x) dout_i, -sum_i q'_i delta_i inv_i).  No GPU: this pins the derivation, the GPU tests
pin the kernels."""
import pytest
import torch

from oracle import performer_oracle as po


def _manual_backward(q, k, v, P, dout, chunk, eps_f=1e-4, eps_c=1e-6):
    B, H, N, d = q.shape
    m = P.shape[0]
    c, r = d ** -0.25, m ** -0.5
    qf, kf = po.softmax_kernel(q, P, True, eps_f), po.softmax_kernel(k, P, False, eps_f)
    out = po.causal_linear_attention(qf, kf, v, eps_c)
    den = (qf * (kf.cumsum(-2) + eps_c)).sum(-1)
    delta, inv = (dout * out).sum(-1), 1.0 / den
    dqf, dkf, dv = torch.zeros_like(qf), torch.zeros_like(kf), torch.zeros_like(v)
    starts = list(range(0, N, chunk))
    # prefix states, forward order
    S = q.new_zeros(B, H, m, d)
    ksum = q.new_zeros(B, H, m) + eps_c
    pre = []
    for s0 in starts:
        pre.append((S.clone(), ksum.clone()))
        kc, vc = kf[:, :, s0:s0 + chunk], v[:, :, s0:s0 + chunk]
        S = S + torch.einsum("bhjm,bhje->bhme", kc, vc)
        ksum = ksum + kc.sum(-2)
    # suffix states, backward order
    R = q.new_zeros(B, H, m, d)
    Rden = q.new_zeros(B, H, m)
    for t in reversed(range(len(starts))):
        s0 = starts[t]
        sl = slice(s0, s0 + chunk)
        qc, kc, vc, dc = qf[:, :, sl], kf[:, :, sl], v[:, :, sl], dout[:, :, sl]
        dl, iv = delta[:, :, sl], inv[:, :, sl]
        St, kst = pre[t]
        Bm = (torch.einsum("bhie,bhje->bhij", dc, vc) - dl.unsqueeze(-1)).tril()                 # rows i, columns j <= i
        dqf[:, :, sl] = iv.unsqueeze(-1) * (torch.einsum("bhie,bhme->bhim", dc, St) - dl.unsqueeze(-1) * kst.unsqueeze(-2)
                                            + torch.einsum("bhij,bhjm->bhim", Bm, kc))
        BT = ((torch.einsum("bhje,bhie->bhji", vc, dc) - dl.unsqueeze(-2)) * iv.unsqueeze(-2)).triu()   # rows j, columns i >= j
        dkf[:, :, sl] = torch.einsum("bhje,bhme->bhjm", vc, R) + Rden.unsqueeze(-2) + torch.einsum("bhji,bhim->bhjm", BT, qc)
        AT = (torch.einsum("bhjm,bhim->bhji", kc, qc) * iv.unsqueeze(-2)).triu()
        dv[:, :, sl] = torch.einsum("bhjm,bhme->bhje", kc, R) + torch.einsum("bhji,bhie->bhje", AT, dc)
        R = R + torch.einsum("bhim,bhie->bhme", qc * iv.unsqueeze(-1), dc)
        Rden = Rden - (qc * (dl * iv).unsqueeze(-1)).sum(-2)
    # feature-map backward
    def fm_bwd(x, feat, dfeat, is_query):
        dD = dfeat * (feat - r * eps_f)
        s = dD.sum(-1, keepdim=True)
        dx = c * torch.einsum("bhnm,md->bhnd", dD, P) - c * c * s * x
        dash = c * torch.einsum("bhnd,md->bhnm", x, P)
        if is_query:
            am = dash.argmax(-1)
            dx = dx - c * s * P[am]
        else:
            flat = int(dash.argmax())
            bb, hh, nn, ff = [int(t) for t in torch.unravel_index(torch.tensor(flat), dash.shape)]
            dx[bb, hh, nn] -= c * s.sum() * P[ff]
        return dx
    return fm_bwd(q, qf, dqf, True), fm_bwd(k, kf, dkf, False), dv, dqf, dkf


@pytest.mark.parametrize("B,H,N,d,m,chunk", [(1, 2, 37, 8, 12, 16), (2, 1, 64, 16, 20, 16), (1, 1, 5, 4, 3, 8)])
def test_chunked_backward_formulas_equal_autograd(B, H, N, d, m, chunk):
    g = torch.Generator().manual_seed(N * m)
    q, k, v = (torch.randn(B, H, N, d, generator=g, dtype=torch.float64).requires_grad_(True) for _ in range(3))
    P = torch.randn(m, d, generator=g, dtype=torch.float64)
    dout = torch.randn(B, H, N, d, generator=g, dtype=torch.float64)
    qf, kf = po.softmax_kernel(q, P, True), po.softmax_kernel(k, P, False)
    qf.retain_grad(); kf.retain_grad()
    out = po.causal_linear_attention(qf, kf, v)
    (out * dout).sum().backward()
    with torch.no_grad():
        dq, dk, dv, dqf, dkf = _manual_backward(q.detach(), k.detach(), v.detach(), P, dout, chunk)
    for name, got, want in (("dq'", dqf, qf.grad), ("dk'", dkf, kf.grad), ("dv", dv, v.grad), ("dq", dq, q.grad),
                            ("dk", dk, k.grad)):
        err = float((got - want).abs().max())
        assert err <= 1e-10 * max(1.0, float(want.abs().max())), f"{name}: {err:.3e}"
