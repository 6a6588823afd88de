"""GPU parity tests of the Performer path: every kernel through the C ABI against the CPU oracle
(oracle/performer_oracle.py) on the same seeded inputs, then the whole network (forward, loss, every parameter
gradient).  fp32 path: tolerance 1e-4 (north_star); bf16 tensor-core path: stated per test."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import performer_oracle as po

pytestmark = pytest.mark.gpu


def _mods():
    from synthanatomy_b200 import ops, pf_ops
    return ops, pf_ops


def _close(got, want, tol=1e-4, what=""):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    scale = max(1.0, float(want.abs().max()))
    err = float((got - want).abs().max())
    assert err <= tol * scale, f"{what}: max abs err {err:.3e} > {tol:.0e} * {scale:.3g}"


# ------------------------------------------------------------------------------------------------ dense layers
@pytest.mark.parametrize("m,n,k", [(130, 70, 50), (64, 64, 16), (257, 129, 100)])
def test_gemm_nt_epilogues_fp32(m, n, k):
    ops, pf = _mods()
    g = torch.Generator().manual_seed(m + n)
    a, b = torch.randn(m, k, generator=g), torch.randn(n, k, generator=g) * 0.2
    bias, resid = torch.randn(n, generator=g), torch.randn(m, n, generator=g)
    w, s = torch.randn(m, n, generator=g), torch.tensor([0.37])
    A, Bm = a.cuda(), b.cuda()
    out = torch.empty(m, n, device="cuda")
    pf.gemm_nt(A, Bm, out_f32=out)
    assert ops.last_path() == 1
    _close(out, a @ b.t(), 1e-5, "plain")
    # bias + GELU forward
    pre, h = torch.empty(m, n, device="cuda"), torch.empty(m, n, device="cuda")
    pf.gemm_nt(A, Bm, bias=bias.cuda(), act=pf.SA_ACT_GELU_FWD, pre=pre, out_act=h)
    u = a @ b.t() + bias
    _close(pre, u, 1e-5, "pre"); _close(h, F.gelu(u), 1e-5, "gelu")
    # ReZero: resid + g * (v + bias), written to a new tensor and in place
    o2 = torch.empty(m, n, device="cuda")
    r = resid.cuda()
    pf.gemm_nt(A, Bm, bias=bias.cuda(), scale_dev=s.cuda(), resid=r, out_f32=o2)
    _close(o2, resid + 0.37 * u, 1e-5, "rezero")
    pf.gemm_nt(A, Bm, bias=bias.cuda(), scale_dev=s.cuda(), resid=r, out_f32=r)
    _close(r, resid + 0.37 * u, 1e-5, "rezero in place")
    # backward-style: dot with a tensor, scale, GELU'
    dot = torch.zeros(1, device="cuda")
    o3 = torch.empty(m, n, device="cuda")
    pf.gemm_nt(A, Bm, dot_with=w.cuda(), dot_out=dot, scale_dev=s.cuda(), scale=2.0, act=pf.SA_ACT_GELU_BWD,
               pre=pre, out_act=o3)
    v = a @ b.t()
    uu = u.clone().requires_grad_(True)
    F.gelu(uu).sum().backward()
    _close(dot, (v * w).sum().view(1), 1e-4, "dot")
    _close(o3, v * 0.74 * uu.grad, 1e-5, "gelu bwd")


def test_gemm_tn_fp32_with_column_slices():
    ops, pf = _mods()
    g = torch.Generator().manual_seed(3)
    big_a, big_b = torch.randn(1000, 200, generator=g), torch.randn(1000, 300, generator=g)
    a, b = big_a[:, 30:100], big_b[:, 64:194]
    d = torch.empty(70, 130, device="cuda")
    A, Bm = big_a.cuda()[:, 30:100], big_b.cuda()[:, 64:194]
    s = torch.tensor([-0.5], device="cuda")
    pf.gemm_tn(A, Bm, d, scale_dev=s, scale=2.0)
    _close(d, -(a.t() @ b), 1e-5, "tn")
    pf.gemm_tn(A, Bm, d, accumulate=True)
    _close(d, torch.zeros(70, 130), 1e-4, "tn accumulate")


def _bf(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("m,n,k,ldo_pad", [(300, 512, 128, 0), (1000, 256, 512, 0), (257, 80, 64, 0), (130, 2049, 192, 0),
                                           (128, 64, 1024, 64), (4099, 1024, 256, 0),
                                           # m >= 1024: CTA-pair kernel (256-row tiles, cta_group::2)
                                           (2500, 2049, 192, 0), (1500, 96, 128, 0), (3000, 512, 512, 64),
                                           (1024, 160, 64, 0), (2000, 512, 1024, 0)])
def test_tcgen05_gemm_nt_bf16(m, n, k, ldo_pad, monkeypatch):
    monkeypatch.setenv("SA_GEMM_PAIR", "1")      # pairs whenever the shape allows (default: only for k >= 1024)
    """bf16 operands / fp32 accumulation on tcgen05; reference = fp32 matmul of the same bf16-rounded operands.
    Tolerance: bf16 output rounding (2^-8 relative) + fp32 accumulation-order slack."""
    ops, pf = _mods()
    g = torch.Generator().manual_seed(m * 7 + n)
    a, b = _bf(torch.randn(m, k, generator=g)), _bf(torch.randn(n, k, generator=g) * 0.1)
    bias, resid = torch.randn(n, generator=g), torch.randn(m, n, generator=g)
    w = _bf(torch.randn(m, n, generator=g))
    s = torch.tensor([0.37])
    A, Bm = a.cuda().bfloat16(), b.cuda().bfloat16()
    ldo = n + ldo_pad

    def buf(dtype):
        return torch.zeros(m, ldo, device="cuda", dtype=dtype)[:, :n]

    v = a @ b.t()
    sc = float(v.abs().max())
    out32 = buf(torch.float32)
    pf.gemm_nt(A, Bm, out_f32=out32)
    assert ops.last_path() == 2, "tcgen05 GEMM was not selected"
    _close(out32, v, 1e-5 * max(1.0, k / 64), "tc plain fp32 out")
    # forward FFN-1 style epilogue
    pre, h = buf(torch.bfloat16), buf(torch.bfloat16)
    pf.gemm_nt(A, Bm, bias=bias.cuda(), act=pf.SA_ACT_GELU_FWD, pre=pre, out_act=h)
    u = v + bias
    torch.testing.assert_close(pre.float().cpu(), u, rtol=2 ** -7, atol=1e-3 * sc)
    torch.testing.assert_close(h.float().cpu(), F.gelu(u), rtol=2 ** -6, atol=2e-3 * sc)
    # ReZero residual epilogue, fp32 stream updated in place + bf16 copy
    r = buf(torch.float32); r.copy_(resid.cuda())
    xb = buf(torch.bfloat16)
    pf.gemm_nt(A, Bm, bias=bias.cuda(), scale_dev=s.cuda(), resid=r, out_f32=r, out_act=xb)
    _close(r, resid + 0.37 * u, 1e-5 * max(1.0, k / 64), "tc rezero")
    torch.testing.assert_close(xb.float().cpu(), resid + 0.37 * u, rtol=2 ** -7, atol=1e-3 * sc)
    # backward-style epilogue
    dot = torch.zeros(1, device="cuda")
    o3 = buf(torch.bfloat16)
    wd = buf(torch.bfloat16); wd.copy_(w.cuda())
    pf.gemm_nt(A, Bm, dot_with=wd, dot_out=dot, scale_dev=s.cuda(), scale=2.0, act=pf.SA_ACT_GELU_BWD, pre=pre, out_act=o3)
    pu = pre.float().cpu().clone().requires_grad_(True)
    F.gelu(pu).sum().backward()
    assert abs(float(dot) - float((v * w).sum())) <= 1e-3 * float((v * w).abs().sum())
    torch.testing.assert_close(o3.float().cpu(), v * 0.74 * pu.grad, rtol=2 ** -6, atol=2e-3 * sc)


@pytest.mark.parametrize("m,na,nb", [(1000, 128, 256), (4096, 512, 1024), (777, 70, 130), (20000, 2049, 512), (64, 128, 64)])
def test_tcgen05_gemm_tn_bf16(m, na, nb):
    ops, pf = _mods()
    g = torch.Generator().manual_seed(m + na)
    lda, ldb = ((na + 7) // 8) * 8 + 8, ((nb + 7) // 8) * 8
    a, b = _bf(torch.randn(m, na, generator=g)), _bf(torch.randn(m, nb, generator=g))
    A = torch.zeros(m, lda, device="cuda", dtype=torch.bfloat16); A[:, 8:8 + na] = a.cuda()
    Bm = torch.zeros(m, ldb, device="cuda", dtype=torch.bfloat16); Bm[:, :nb] = b.cuda()
    d = torch.empty(na, nb, device="cuda")
    s = torch.tensor([-0.5], device="cuda")
    pf.gemm_tn(A[:, 8:8 + na], Bm[:, :nb], d, scale_dev=s, scale=2.0)
    assert ops.last_path() == 2, "tcgen05 wgrad GEMM was not selected"
    want = -(a.t() @ b)
    _close(d, want, 2e-5 * max(1.0, m / 1000), "tc tn")
    pf.gemm_tn(A[:, 8:8 + na], Bm[:, :nb], d, accumulate=True)
    assert float(d.abs().max()) <= 1e-4 * float(want.abs().max()) * max(1.0, m / 1000)
    # the same call with the column sums of A (the bias gradient) as one more product of the staged tiles
    cs = torch.full((na,), 7.0, device="cuda")
    d2 = torch.empty(na, nb, device="cuda")
    pf.gemm_tn(A[:, 8:8 + na], Bm[:, :nb], d2, scale_dev=s, scale=2.0, colsum=cs)
    assert ops.last_path() == 2
    _close(d2, want, 2e-5 * max(1.0, m / 1000), "tc tn + colsum")
    torch.testing.assert_close(cs.cpu(), a.sum(0), rtol=1e-5, atol=1e-5 * float(a.abs().sum(0).max()))


def test_gemm_tn_colsum_cuda_core_path():
    ops, pf = _mods()
    g = torch.Generator().manual_seed(11)
    a, b = torch.randn(500, 72, generator=g), torch.randn(500, 40, generator=g)
    d, cs = torch.empty(72, 40, device="cuda"), torch.empty(72, device="cuda")
    pf.gemm_tn(a.cuda(), b.cuda(), d, colsum=cs)
    _close(d, a.t() @ b, 1e-5, "tn")
    torch.testing.assert_close(cs.cpu(), a.sum(0), rtol=1e-5, atol=1e-4)


# ------------------------------------------------------------------------------------------------ FAVOR+
def _heads_to_rows(t):          # [B, H, N, d] -> [B*N, H*d]
    B, H, N, d = t.shape
    return t.permute(0, 2, 1, 3).reshape(B * N, H * d).contiguous()


def _rows_to_heads(t, B, H):    # [B*N, H*d] -> [B, H, N, d]
    M, C = t.shape
    return t.view(B, M // B, H, C // H).permute(0, 2, 1, 3)


@pytest.mark.parametrize("B,H,N,m", [(2, 2, 150, 266), (1, 3, 37, 40)])
def test_favor_featmap_fwd_bwd_fp32(B, H, N, m):
    ops, pf = _mods()
    g = torch.Generator().manual_seed(N)
    d, mp = 64, ((m + 15) // 16) * 16
    q = torch.randn(B, H, N, d, generator=g).requires_grad_(True)
    k = torch.randn(B, H, N, d, generator=g).requires_grad_(True)
    P = po.gaussian_orthogonal_random_matrix(m, d, generator=g)
    wq, wk = torch.randn(B, H, N, m, generator=g), torch.randn(B, H, N, m, generator=g)
    qf, kf = po.softmax_kernel(q, P, True), po.softmax_kernel(k, P, False)
    ((qf * wq).sum() + (kf * wk).sum()).backward()
    kmax_ref = float((k.detach() * d ** -0.25 @ P.t()).max())

    ld = 2 * H * d + 8                      # q block | k block | padding: exercises the leading dimension
    buf = torch.zeros(B * N, ld)
    buf[:, :H * d] = _heads_to_rows(q.detach()); buf[:, H * d:2 * H * d] = _heads_to_rows(k.detach())
    buf = buf.cuda()
    fd = pf.favor_desc(B, N, H, d, m, mp, ld, torch.float32)
    kmax = torch.zeros(1, dtype=torch.int64, device="cuda")
    pf.favor_kmax(fd, buf, H * d, P.cuda(), kmax)
    packed = int(kmax.item()) & 0xFFFFFFFFFFFFFFFF
    bits = packed >> 32
    bits = bits ^ 0x80000000 if bits & 0x80000000 else (~bits) & 0xFFFFFFFF
    assert abs(np.frombuffer(np.uint32(bits).tobytes(), dtype=np.float32)[0] - kmax_ref) <= 1e-5 * abs(kmax_ref)
    QF = torch.empty(B, H, N, mp, device="cuda"); KF = torch.empty(B, H, N, mp, device="cuda")
    argq = torch.empty(B, H, N, dtype=torch.int32, device="cuda")
    pf.favor_featmap_fwd(fd, buf, 0, P.cuda(), True, None, 1e-4, QF, argq)
    pf.favor_featmap_fwd(fd, buf, H * d, P.cuda(), False, kmax, 1e-4, KF, None)
    _close(QF[..., :m], qf, 1e-5, "q features"); _close(KF[..., :m], kf, 1e-5, "k features")
    assert float(QF[..., m:].abs().max()) == 0.0 if mp > m else True
    # backward
    dQF = torch.zeros(B, H, N, mp); dQF[..., :m] = wq
    dKF = torch.zeros(B, H, N, mp); dKF[..., :m] = wk
    dbuf = torch.zeros(B * N, ld, device="cuda")
    gsum = torch.zeros(1, device="cuda")
    pf.favor_featmap_bwd(fd, buf, 0, P.cuda(), True, 1e-4, QF, dQF.cuda(), argq, dbuf, 0, None)
    pf.favor_featmap_bwd(fd, buf, H * d, P.cuda(), False, 1e-4, KF, dKF.cuda(), None, dbuf, H * d, gsum)
    pf.favor_kmax_fixup(fd, P.cuda(), kmax, gsum, dbuf, H * d)
    _close(_rows_to_heads(dbuf[:, :H * d].cpu(), B, H), q.grad, 1e-4, "dq")
    _close(_rows_to_heads(dbuf[:, H * d:2 * H * d].cpu(), B, H), k.grad, 1e-4, "dk")


@pytest.mark.parametrize("B,H,N,m", [(2, 2, 150, 266), (1, 1, 64, 266), (1, 2, 7, 30)])
def test_favor_scan_fwd_bwd_fp32(B, H, N, m):
    ops, pf = _mods()
    g = torch.Generator().manual_seed(N + m)
    d, mp = 64, ((m + 15) // 16) * 16
    qf = (torch.rand(B, H, N, m, generator=g) * 0.1 + 1e-3).requires_grad_(True)
    kf = (torch.rand(B, H, N, m, generator=g) * 0.1 + 1e-3).requires_grad_(True)
    v = torch.randn(B, H, N, d, generator=g).requires_grad_(True)
    w = torch.randn(B, H, N, d, generator=g)
    out = po.causal_linear_attention(qf, kf, v)
    (out * w).sum().backward()

    QF = torch.zeros(B, H, N, mp); QF[..., :m] = qf.detach()
    KF = torch.zeros(B, H, N, mp); KF[..., :m] = kf.detach()
    QF, KF = QF.cuda(), KF.cuda()
    ld = H * d + 16
    vbuf = torch.zeros(B * N, ld); vbuf[:, 16:] = _heads_to_rows(v.detach()); vbuf = vbuf.cuda()
    fd = pf.favor_desc(B, N, H, d, m, mp, ld, torch.float32)
    ws = torch.empty(pf.favor_scan_workspace(fd, True), dtype=torch.uint8, device="cuda")
    O = torch.zeros(B * N, H * d + 64, device="cuda")
    den = torch.empty(B, H, N, device="cuda")
    pf.favor_scan_fwd(fd, QF, KF, vbuf, 16, 1e-6, O, 64, den, ws)
    _close(_rows_to_heads(O[:, 64:].cpu(), B, H), out, 1e-4, "scan out")
    dO = torch.zeros(B * N, H * d + 64); dO[:, 64:] = _heads_to_rows(w); dO = dO.cuda()
    dQF, dKF = torch.empty_like(QF), torch.empty_like(KF)
    dv = torch.zeros(B * N, ld, device="cuda")
    pf.favor_scan_bwd(fd, QF, KF, vbuf, 16, 1e-6, O, dO, 64, den, dQF, dKF, dv, 16, ws)
    _close(dQF[..., :m], qf.grad, 1e-4, "dq'"); _close(dKF[..., :m], kf.grad, 1e-4, "dk'")
    _close(_rows_to_heads(dv[:, 16:].cpu(), B, H), v.grad, 1e-4, "dv")


# ------------------------------------------------------------------------------------------------ local heads
@pytest.mark.parametrize("B,H,N,W,rot", [(2, 2, 150, 40, True), (1, 3, 200, 64, False), (1, 1, 19, 20, True),
                                         (1, 2, 130, 7, True)])
def test_local_attention_fwd_bwd_fp32(B, H, N, W, rot):
    ops, pf = _mods()
    g = torch.Generator().manual_seed(N + W)
    d = 64
    q, k, v = (torch.randn(B, H, N, d, generator=g).requires_grad_(True) for _ in range(3))
    w = torch.randn(B, H, N, d, generator=g)
    out = po.local_attention(q, k, v, W, "rotary" if rot else "none")
    (out * w).sum().backward()
    inner = H * d
    buf = torch.cat([_heads_to_rows(t.detach()) for t in (q, k, v)], dim=1).cuda()
    inv_freq = (1.0 / (10000 ** (torch.arange(0, d, 2).float() / d))).cuda() if rot else None
    ldsc = pf.local_desc(B, N, H, d, W, 3 * inner, inner, torch.float32)
    O = torch.empty(B * N, inner, device="cuda")
    lse = torch.empty(B, H, N, device="cuda")
    pf.local_attn_fwd(ldsc, buf, 0, inner, 2 * inner, inv_freq, O, 0, lse)
    _close(_rows_to_heads(O.cpu(), B, H), out, 1e-4, "local out")
    dbuf = torch.zeros_like(buf)
    pf.local_attn_bwd(ldsc, buf, 0, inner, 2 * inner, inv_freq, O, _heads_to_rows(w).cuda(), 0, lse, dbuf)
    for i, (name, t) in enumerate((("dq", q), ("dk", k), ("dv", v))):
        _close(_rows_to_heads(dbuf[:, i * inner:(i + 1) * inner].cpu(), B, H), t.grad, 1e-4, name)


@pytest.mark.parametrize("B,H,N,W", [(2, 2, 300, 40), (1, 3, 1000, 420), (1, 1, 150, 64), (2, 1, 130, 7), (1, 2, 1400, 420)])
def test_tcgen05_local_attention_bf16(B, H, N, W):
    """tcgen05 flash-style kernels (bf16 operands, bf16 P / dS re-staging, fp32 accumulation) vs the oracle on the same
    bf16-rounded inputs; tolerance 2e-2 of the tensor's max (bf16 rounding of probabilities), not the fp32 claim."""
    ops, pf = _mods()
    g = torch.Generator().manual_seed(N + W)
    d = 64
    q, k, v = (_bf(torch.randn(B, H, N, d, generator=g)).requires_grad_(True) for _ in range(3))
    w = _bf(torch.randn(B, H, N, d, generator=g))
    out = po.local_attention(q, k, v, W, "none")
    (out * w).sum().backward()
    inner = H * d
    buf = torch.cat([_heads_to_rows(t.detach()) for t in (q, k, v)], dim=1).cuda().bfloat16()
    ldsc = pf.local_desc(B, N, H, d, W, 3 * inner, inner, torch.bfloat16)
    O = torch.zeros(B * N, inner, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device="cuda")
    pf.local_attn_fwd(ldsc, buf, 0, inner, 2 * inner, None, O, 0, lse)
    assert ops.last_path() == 2, "tcgen05 local attention was not selected"
    _close(_rows_to_heads(O.float().cpu(), B, H), out, 2e-2, "tc local out")
    dbuf = torch.zeros_like(buf)
    pf.local_attn_bwd(ldsc, buf, 0, inner, 2 * inner, None, O, _heads_to_rows(w).cuda().bfloat16(), 0, lse, dbuf)
    assert ops.last_path() == 2
    for i, (name, t) in enumerate((("dq", q), ("dk", k), ("dv", v))):
        got = _rows_to_heads(dbuf[:, i * inner:(i + 1) * inner].float().cpu(), B, H)
        scale = float(t.grad.abs().max())
        err = float((got - t.grad).abs().max())
        assert err <= 2e-2 * scale, f"tc local {name}: {err:.3e} vs max {scale:.3e}"
    # the CUDA-core kernel on the same bf16 buffers agrees too (same masks, same window arithmetic)
    ops.set_force_simt(True)
    try:
        O2 = torch.zeros_like(O); lse2 = torch.empty_like(lse)
        pf.local_attn_fwd(ldsc, buf, 0, inner, 2 * inner, None, O2, 0, lse2)
        assert ops.last_path() == 1
    finally:
        ops.set_force_simt(False)
    _close(lse, lse2, 1e-3, "lse tc vs simt")


@pytest.mark.parametrize("B,H,N,W,simt", [(2, 2, 300, 40, False), (1, 2, 1400, 420, False), (1, 1, 150, 64, False),
                                          (2, 1, 130, 7, False), (1, 2, 300, 40, True)])
def test_local_attention_rotary_transpose_in_the_backward_epilogue(B, H, N, W, simt):
    """q / k rotated in place, forward on the rotated buffers, sa_local_attn_bwd_rot: dq / dk leave the tcgen05 kernels
    through the transpose of the rotation (registers, fp32).  Against the oracle's rotary local attention on the same
    bf16-rounded inputs; and the CUDA-core branch of the same entry point (rotation as its own pass)."""
    ops, pf = _mods()
    g = torch.Generator().manual_seed(N + W + 1)
    d = 64
    q, k, v = (_bf(torch.randn(B, H, N, d, generator=g)).requires_grad_(True) for _ in range(3))
    w = _bf(torch.randn(B, H, N, d, generator=g))
    out = po.local_attention(q, k, v, W, "rotary")
    (out * w).sum().backward()
    inner = H * d
    buf = torch.cat([_heads_to_rows(t.detach()) for t in (q, k, v)], dim=1).cuda().bfloat16()
    inv_freq = (1.0 / (10000 ** (torch.arange(0, d, 2).float() / d))).cuda()
    table = pf.rotary_table(inv_freq, N, d)
    ang = torch.arange(N, dtype=torch.float32)[:, None] * inv_freq.cpu()[None, :]
    _close(table[..., 0].cpu(), torch.cos(ang), 1e-5, "cos table"); _close(table[..., 1].cpu(), torch.sin(ang), 1e-5, "sin table")
    ldsc = pf.local_desc(B, N, H, d, W, 3 * inner, inner, torch.bfloat16)
    O = torch.zeros(B * N, inner, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device="cuda")
    ops.set_force_simt(simt)
    try:
        pf.rotary_qk(buf, 0, inner, B, N, H, d, inv_freq, False)
        pf.local_attn_fwd(ldsc, buf, 0, inner, 2 * inner, None, O, 0, lse)
        _close(_rows_to_heads(O.float().cpu(), B, H), out, 3e-2, "rotary local out")
        dbuf = torch.zeros_like(buf)
        dO = _heads_to_rows(w).cuda().bfloat16()
        pf.local_attn_bwd_rot(ldsc, buf, 0, inner, 2 * inner, inv_freq, table, O, dO, 0, lse, dbuf)
        assert ops.last_path() == (1 if simt else 2)
        # the two-pass form on the same buffers: gradients of the rotated q / k, then the in-place transpose
        dref = torch.zeros_like(buf)
        pf.local_attn_bwd(ldsc, buf, 0, inner, 2 * inner, None, O, dO, 0, lse, dref)
        pf.rotary_qk(dref, 0, inner, B, N, H, d, inv_freq, True)
    finally:
        ops.set_force_simt(False)
    for i, (name, t) in enumerate((("dq", q), ("dk", k), ("dv", v))):
        got = _rows_to_heads(dbuf[:, i * inner:(i + 1) * inner].float().cpu(), B, H)
        two = _rows_to_heads(dref[:, i * inner:(i + 1) * inner].float().cpu(), B, H)
        scale = float(t.grad.abs().max())
        err = float((got - t.grad).abs().max())
        assert err <= 3e-2 * scale, f"rot local {name}: {err:.3e} vs max {scale:.3e}"
        # one bf16 rounding instead of two: the fused form differs from the two-pass form by bf16 ulps only
        assert float((got - two).abs().max()) <= 2.0 ** -6 * scale, name
    assert torch.equal(dbuf[:, 2 * inner:], dref[:, 2 * inner:]), "dv does not depend on where the rotation is undone"


def test_rotary_inplace_matches_oracle():
    ops, pf = _mods()
    g = torch.Generator().manual_seed(2)
    B, H, N, d = 2, 3, 77, 64
    q, k = torch.randn(B, H, N, d, generator=g), torch.randn(B, H, N, d, generator=g)
    qr, kr = po.apply_rotary_pos_emb(q, k, po.sinusoidal_embeddings(N, d)[None, None])
    buf = torch.cat([torch.zeros(B * N, 64), _heads_to_rows(q)], dim=1).cuda()
    inv_freq = (1.0 / (10000 ** (torch.arange(0, d, 2).float() / d))).cuda()
    pf.rotary(buf, 64, B, N, H, d, inv_freq, False)
    _close(_rows_to_heads(buf[:, 64:].cpu(), B, H), qr, 1e-5, "rotary")
    assert float(buf[:, :64].abs().max()) == 0.0
    pf.rotary(buf, 64, B, N, H, d, inv_freq, True)
    _close(_rows_to_heads(buf[:, 64:].cpu(), B, H), q, 1e-5, "rotary inverse")


# ------------------------------------------------------------------------------------------------ ends
def test_layernorm_ce_embed_fp32():
    ops, pf = _mods()
    g = torch.Generator().manual_seed(9)
    rows, dim, V = 300, 96, 77
    x = torch.randn(rows, dim, generator=g).requires_grad_(True)
    w = torch.randn(dim, generator=g).requires_grad_(True); b = torch.randn(dim, generator=g).requires_grad_(True)
    gy = torch.randn(rows, dim, generator=g)
    y = F.layer_norm(x, (dim,), w, b, 1e-5); y.backward(gy)
    Y = torch.empty(rows, dim, device="cuda"); mean = torch.empty(rows, device="cuda"); rstd = torch.empty(rows, device="cuda")
    pf.layernorm_fwd(x.detach().cuda(), w.detach().cuda(), b.detach().cuda(), 1e-5, Y, None, mean, rstd)
    _close(Y, y, 1e-5, "LN")
    dx = torch.empty(rows, dim, device="cuda"); dw = torch.zeros(dim, device="cuda"); db = torch.zeros(dim, device="cuda")
    pf.layernorm_bwd(gy.cuda(), x.detach().cuda(), w.detach().cuda(), mean, rstd, dx, dw, db)
    _close(dx, x.grad, 1e-5, "LN dx"); _close(dw, w.grad, 1e-4, "LN dw"); _close(db, b.grad, 1e-4, "LN db")
    # cross-entropy (sum of per-row losses, gradient of the mean)
    logits = (torch.randn(rows, V, generator=g) * 3).requires_grad_(True)
    tgt = torch.randint(0, V, (rows,), generator=g)
    loss = F.cross_entropy(logits, tgt, reduction="mean"); loss.backward()
    ls = torch.zeros(1, device="cuda"); dl = torch.empty(rows, V, device="cuda")
    pf.ce_fwd_bwd(logits.detach().cuda(), tgt.cuda(), 1.0 / rows, None, ls, dl)
    _close(ls / rows, loss.view(1), 1e-5, "CE"); _close(dl, logits.grad, 1e-6, "CE grad")
    # embeddings
    Bn, N, nt = 2, 11, 13
    tok = torch.randint(0, nt, (Bn, N), generator=g)
    sp = torch.randint(0, 5, (2, N), generator=g).to(torch.int32); sp[:, 0] = -1
    tw = torch.randn(nt, dim, generator=g, requires_grad=True); pw = torch.randn(N + 3, dim, generator=g, requires_grad=True)
    s0 = torch.randn(5, dim, generator=g, requires_grad=True); s1 = torch.randn(5, dim, generator=g, requires_grad=True)
    ref = tw[tok] + pw[:N]
    for s_w, row in ((s0, sp[0]), (s1, sp[1])):
        ref = ref + torch.where(row[:, None] >= 0, s_w[row.clamp(min=0).long()], torch.zeros(()))
    gx = torch.randn(Bn, N, dim, generator=g); ref.backward(gx)
    X = torch.empty(Bn * N, dim, device="cuda")
    pf.embed_fwd(tok.cuda(), sp.cuda(), tw.detach().cuda(), [s0.detach().cuda(), s1.detach().cuda()], pw.detach().cuda(), X, None)
    _close(X.view(Bn, N, dim), ref, 1e-6, "embed")
    dtw = torch.zeros(nt, dim, device="cuda"); dpw = torch.zeros(N + 3, dim, device="cuda")
    ds = [torch.zeros(5, dim, device="cuda") for _ in range(2)]
    pf.embed_bwd(gx.cuda().view(Bn * N, dim).contiguous(), tok.cuda(), sp.cuda(), dtw, ds, dpw)
    _close(dtw, tw.grad, 1e-5, "d tok"); _close(dpw, pw.grad, 1e-5, "d pos")
    _close(ds[0], s0.grad, 1e-5, "d sp0"); _close(ds[1], s1.grad, 1e-5, "d sp1")


# ------------------------------------------------------------------------------------------------ whole network
def _build(cfg_kw, grid, seed, compute_dtype=None):
    from synthanatomy_b200.networks.transformers import Ordering, Performer
    order = Ordering("raster_scan", 3, (1, *grid), (False,) * 3, ((2, 0, 1),), ((0, 1),), ("rotate_90", "transpose"))
    n = int(np.prod(grid))
    cfg = po.PerformerConfig(max_seq_len=n + 1, spatial_shape=tuple(grid), **cfg_kw)
    sd = po.init_state_dict(cfg, seed)
    # a ReZero gate of 1e-3 hides everything behind the residual: open the gates so that every kernel matters
    for i in range(cfg.depth):
        sd[po.layer_prefix(i) + "0.g"] = torch.tensor(0.7); sd[po.layer_prefix(i) + "1.g"] = torch.tensor(-0.4)
    net = Performer(num_tokens=cfg.num_tokens, max_seq_len=n + 1, dim=cfg.dim, depth=cfg.depth, heads=cfg.heads,
                    ordering=order, dim_head=cfg.dim_head, local_attn_heads=cfg.local_attn_heads,
                    local_window_size=cfg.local_window_size, feature_redraw_interval=1, use_rezero=True,
                    spatial_position_emb="absolute", spatial_shape=tuple(grid), compute_dtype=compute_dtype)
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected
    assert all(("proj_updater" in k) or k.endswith(("inv_freq", "spatial_indices_sequence")) for k in missing), missing
    net.fix_projection_matrices_()
    seqs = [torch.from_numpy(s.copy()) for s in po.spatial_index_sequences(grid, order.get_sequence_ordering())]
    g = torch.Generator().manual_seed(seed + 1)
    q = torch.randint(0, cfg.num_tokens - 1, (2, *grid), generator=g)
    x_in, y = po.prepare_batch(q.numpy(), order.get_sequence_ordering(), cfg.num_tokens - 1)
    return cfg, sd, net, seqs, torch.from_numpy(x_in), torch.from_numpy(y)


CASES = {
    "tiny": (dict(num_tokens=65, dim=128, depth=2, heads=4, dim_head=64, local_attn_heads=2, local_window_size=20), (4, 5, 6)),
    "ragged_window": (dict(num_tokens=33, dim=64, depth=1, heads=2, dim_head=64, local_attn_heads=1, local_window_size=33),
                      (3, 7, 5)),
    "global_only": (dict(num_tokens=33, dim=64, depth=1, heads=2, dim_head=64, local_attn_heads=0, local_window_size=16),
                    (2, 5, 7)),
    "readme_slice": (dict(num_tokens=2049, dim=512, depth=2, heads=16, dim_head=64, local_attn_heads=8,
                          local_window_size=420), (10, 14, 10)),
}


@pytest.mark.parametrize("name", list(CASES))
def test_performer_forward_backward_matches_oracle_fp32(name):
    from synthanatomy_b200.losses import CELoss
    kw, grid = CASES[name]
    cfg, sd, net, seqs, x_in, y = _build(kw, grid, 11)
    loss_ref, grads_ref, logits_ref = po.train_step_grads(sd, cfg, x_in, y, seqs)
    net = net.cuda().train()
    logits = net(x_in.cuda())
    _close(logits, logits_ref, 1e-4, "logits")
    loss = CELoss()(logits.transpose(1, 2), y.cuda())
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * max(1.0, abs(float(loss_ref)))
    loss.backward()
    named = dict(net.named_parameters())
    worst = 0.0
    for k, gref in grads_ref.items():
        got = named[k].grad
        assert got is not None, k
        scale = max(float(gref.abs().max()), 1e-3)
        err = float((got.cpu() - gref).abs().max()) / scale
        worst = max(worst, err)
        assert err <= 2e-4, f"grad {k}: max err relative to max |grad| {err:.3e}"
    # eval-mode API used by the sampling loop: same logits, no saved state
    net.eval()
    with torch.no_grad():
        _close(net(x_in.cuda()[:, :17]), po.forward(sd, cfg, x_in[:, :17], seqs), 1e-4, "prefix forward")
        enc = net(x_in.cuda(), return_encodings=True)
    _close(enc, po.forward(sd, cfg, x_in, seqs, return_encodings=True), 1e-4, "encodings")


def test_performer_bf16_path_tracks_oracle():
    """bf16 operands, fp32 accumulation / residual stream: a stated, looser tolerance (not the 1e-4 parity claim)."""
    from synthanatomy_b200.losses import CELoss
    kw, grid = CASES["readme_slice"]
    cfg, sd, net, seqs, x_in, y = _build(kw, grid, 5, compute_dtype=torch.bfloat16)
    loss_ref, grads_ref, logits_ref = po.train_step_grads(sd, cfg, x_in, y, seqs)
    net = net.cuda().train()
    logits = net(x_in.cuda())
    err = float((logits.cpu() - logits_ref).abs().max()) / float(logits_ref.abs().max())
    assert err < 5e-2, f"bf16 logits relative error {err}"
    loss = CELoss()(logits.transpose(1, 2), y.cuda())
    assert abs(float(loss) - float(loss_ref)) < 2e-2 * abs(float(loss_ref))
    loss.backward()
    named = dict(net.named_parameters())
    for k in ("to_out.weight", "performer.net.layers.0.1.fn.fn.w1.weight", "performer.net.layers.0.0.fn.to_q.weight",
              "performer.net.layers.1.0.fn.to_v.weight", "token_emb.weight"):
        gref = grads_ref[k]
        cos = F.cosine_similarity(named[k].grad.cpu().flatten(), gref.flatten(), dim=0)
        assert float(cos) > 0.98, f"bf16 grad {k}: cosine {float(cos)}"


def test_sampling_loop_and_adam_step():
    from synthanatomy_b200.optim import Adam
    kw, grid = CASES["global_only"][0], (2, 3, 2)
    cfg, sd, net, seqs, x_in, y = _build(dict(kw, local_attn_heads=1, local_window_size=4), grid, 3)
    net = net.cuda()
    prefix = torch.full((2, 1), cfg.num_tokens - 1, dtype=torch.long, device="cuda")
    out = net.sample(prefix, sample=False)
    assert tuple(out.shape) == (2, *grid) and int(out.max()) < cfg.num_tokens
    # greedy sampling == arg-max of the oracle's logits, token by token
    x = torch.full((2, 1), cfg.num_tokens - 1, dtype=torch.long)
    for _ in range(int(np.prod(grid))):
        nxt = po.forward(sd, cfg, x, seqs)[:, -1].argmax(-1, keepdim=True)
        x = torch.cat((x, nxt), dim=1)
    want = x[:, 1:][:, net.ordering.get_revert_sequence_ordering()].reshape(2, *grid)
    assert torch.equal(out.cpu(), want)
    # one Adam step on the Performer parameters against the oracle's Adam
    net.train()
    from synthanatomy_b200.losses import CELoss
    opt = Adam(net.parameters(), lr=1e-3)
    loss = CELoss()(net(x_in.cuda()).transpose(1, 2), y.cuda()); loss.backward(); opt.step()
    _, grads_ref, _ = po.train_step_grads(sd, cfg, x_in, y, seqs)
    k = "performer.net.layers.0.1.fn.fn.w2.weight"
    want_p, _, _ = po.adam_step(sd[k], grads_ref[k], torch.zeros_like(sd[k]), torch.zeros_like(sd[k]), 1, 1e-3)
    # Adam's first step is lr * sign(g) wherever |g| >> eps: compare where the gradient is not tiny
    big = grads_ref[k].abs() > 1e-6
    _close(dict(net.named_parameters())[k].detach().cpu()[big], want_p[big], 1e-4, "adam")


CASES["tiny_depth1"] = (dict(CASES["tiny"][0], depth=1), CASES["tiny"][1])


@pytest.mark.parametrize("name,dtype,tol", [("tiny_depth1", None, 1e-4), ("ragged_window", None, 1e-4),
                                            ("global_only", None, 1e-4), ("tiny_depth1", torch.bfloat16, 5e-2),
                                            ("tiny", None, 2e-2)])
def test_recurrent_decoder_matches_prefix_forward(name, dtype, tol):
    """the recurrent-state decoder (one position per call) against the quantity the reference's
    sampling loop uses: the logits of a forward over the prefix x[:, :t+1] at its last position.

    One attention layer: exact (1e-4 in fp32), including the prefix-dependent global key stabiliser of the FAVOR+
    heads (the carried sums are re-normalised when it moves).  Deeper stacks: the reference itself is not causal --
    extending the prefix moves the stabiliser and thereby (through the +eps of the feature map) the outputs of EARLIER
    positions of a layer, which the next layer's keys / values are made of -- so no O(1)-state evaluation can reproduce
    it bit for bit; the stated tolerance (2e-2 of max |logit| with the ReZero gates opened to 0.7 / -0.4 to amplify the
    effect) bounds that coupling."""
    kw, grid = CASES[name]
    cfg, sd, net, seqs, x_in, y = _build(kw, grid, 21, compute_dtype=dtype)
    net = net.cuda().eval()
    x = x_in.cuda()
    n = x.shape[1]
    dec = net.make_decoder(x.shape[0], n)
    checks = sorted(set([0, 1, 2, 5, kw["local_window_size"], kw["local_window_size"] + 1, 2 * kw["local_window_size"] + 3,
                         n // 2, n - 1]))
    with torch.no_grad():
        for t in range(n):
            lg = dec.step(x[:, t], t)
            if t in checks and t < n:
                ref = net(x[:, :t + 1])[:, -1]
                scale = max(1.0, float(ref.abs().max()))
                err = float((lg - ref).abs().max())
                assert err <= tol * scale, f"position {t}: decoder logits differ from the prefix forward by {err:.3e}"
    if dtype is None and kw["depth"] == 1:      # and from the CPU oracle's prefix forward at the last position
        want = po.forward(sd, cfg, x_in, seqs)[:, -1]
        _close(lg, want, 1e-4, "decoder vs oracle at the last position")


def test_recurrent_sampling_equals_full_forward_sampling():
    kw, grid = CASES["tiny_depth1"]
    cfg, sd, net, seqs, x_in, y = _build(kw, grid, 23)
    net = net.cuda().eval()
    prefix = torch.full((2, 1), cfg.num_tokens - 1, dtype=torch.long, device="cuda")
    a = net.sample(prefix, sample=False, recurrent=True)
    b = net.sample(prefix, sample=False)          # default = the reference's loop of full forwards
    assert tuple(a.shape) == (2, *grid)
    assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------------ wrapper options
def _build_opts(seed, **opts):
    """tiny network with the wrapper options of performer.py:43-67 (fixed spatial code), :136-138 (fixed position code)
    and :183-187 / :248-280 (conditioning)"""
    from synthanatomy_b200.networks.transformers import Ordering, Performer
    kw, grid = CASES["tiny"]
    order = Ordering("raster_scan", 3, (1, *grid), (False,) * 3, ((2, 0, 1),), ((0, 1),), ("rotate_90", "transpose"))
    n = int(np.prod(grid))
    cfg = po.PerformerConfig(max_seq_len=n + 1, spatial_shape=tuple(grid), **kw, **opts)
    sd = po.init_state_dict(cfg, seed)
    for i in range(cfg.depth):
        sd[po.layer_prefix(i) + "0.g"] = torch.tensor(0.7); sd[po.layer_prefix(i) + "1.g"] = torch.tensor(-0.4)
    net = Performer(num_tokens=cfg.num_tokens, max_seq_len=n + 1, dim=cfg.dim, depth=cfg.depth, heads=cfg.heads,
                    ordering=order, dim_head=cfg.dim_head, local_attn_heads=cfg.local_attn_heads,
                    local_window_size=cfg.local_window_size, feature_redraw_interval=1, use_rezero=True,
                    spatial_position_emb=cfg.spatial_position_emb, spatial_shape=tuple(grid),
                    fixed_position_emb=cfg.fixed_position_emb, conditioning_num_tokens=cfg.conditioning_num_tokens,
                    conditioning_type=cfg.conditioning_type)
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(("proj_updater" in k) or k.endswith(("inv_freq", "spatial_indices_sequence", ".emb")) for k in missing), missing
    net.fix_projection_matrices_()
    seqs = [torch.from_numpy(s.copy()) for s in po.spatial_index_sequences(grid, order.get_sequence_ordering())]
    g = torch.Generator().manual_seed(seed + 1)
    q = torch.randint(0, cfg.num_tokens - 1, (2, *grid), generator=g)
    x_in, y = po.prepare_batch(q.numpy(), order.get_sequence_ordering(), cfg.num_tokens - 1)
    return cfg, sd, net, seqs, torch.from_numpy(x_in), torch.from_numpy(y)


@pytest.mark.parametrize("opts", [
    dict(spatial_position_emb="fixed"),
    dict(fixed_position_emb=True),
    dict(conditioning_num_tokens=(5, 3), conditioning_type="bos_replacement"),
    dict(conditioning_num_tokens=(5, 3), conditioning_type="prepending"),
], ids=["fixed_spatial", "fixed_position", "bos_replacement", "prepending"])
def test_wrapper_options_match_oracle_fp32(opts):
    from synthanatomy_b200.losses import CELoss
    cfg, sd, net, seqs, x_in, y = _build_opts(41, **opts)
    conds = None
    if cfg.conditioning_num_tokens:
        g = torch.Generator().manual_seed(7)
        conds = [torch.randint(0, cnt, (2, 1), generator=g) for cnt in cfg.conditioning_num_tokens]
    if cfg.fixed_position_emb or cfg.spatial_position_emb == "fixed":      # the drop-in's own buffers = the formula
        for k, v in net.state_dict().items():
            if k.endswith(".emb") and k in sd:
                torch.testing.assert_close(v, sd[k])
    loss_ref, grads_ref, logits_ref = po.train_step_grads(sd, cfg, x_in, y, seqs, conditionings=conds)
    net = net.cuda().train()
    logits = net(x_in.cuda(), [c.cuda() for c in conds] if conds else None)
    assert tuple(logits.shape) == tuple(logits_ref.shape)
    _close(logits, logits_ref, 1e-4, "logits")
    loss = CELoss()(logits.transpose(1, 2), y.cuda())
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * max(1.0, abs(float(loss_ref)))
    loss.backward()
    named = dict(net.named_parameters())
    for k, gref in grads_ref.items():
        got = named[k].grad
        assert got is not None, k
        err = float((got.cpu() - gref).abs().max()) / max(float(gref.abs().max()), 1e-3)
        assert err <= 2e-4, f"grad {k}: {err:.3e}"


@pytest.mark.parametrize("dtype,m,n,k", [(torch.float32, 130, 70, 50), (torch.bfloat16, 300, 512, 128), (torch.bfloat16, 2500, 2048, 512),
                                         (torch.bfloat16, 257, 80, 64)])
def test_gemm_gelu_derivative_epilogues(dtype, m, n, k):
    """SA_ACT_GELU_FWD_D stores gelu'(u) instead of u, SA_ACT_MUL_PRE multiplies by it: together they equal the
    SA_ACT_GELU_FWD / SA_ACT_GELU_BWD pair (one erf evaluation per element instead of two)."""
    ops, pf = _mods()
    g = torch.Generator().manual_seed(m + k)
    rnd = (lambda *s: _bf(torch.randn(*s, generator=g))) if dtype == torch.bfloat16 else (lambda *s: torch.randn(*s, generator=g))
    a, b, a2, b2 = rnd(m, k), rnd(n, k) * 0.1, rnd(m, 64), rnd(n, 64) * 0.1
    bias, w = torch.randn(n, generator=g), rnd(m, n)
    s = torch.tensor([0.37])
    A, Bm, A2, B2 = (t.cuda().to(dtype) for t in (a, b, a2, b2))
    tol = 1e-5 if dtype == torch.float32 else 2 ** -6
    u = a @ b.t() + bias
    uu = u.clone().requires_grad_(True)
    F.gelu(uu).sum().backward()
    d, h = torch.empty(m, n, device="cuda", dtype=dtype), torch.empty(m, n, device="cuda", dtype=dtype)
    pf.gemm_nt(A, Bm, bias=bias.cuda(), act=pf.SA_ACT_GELU_FWD_D, pre=d, out_act=h)
    torch.testing.assert_close(h.float().cpu(), F.gelu(u), rtol=tol, atol=tol * float(u.abs().max()))
    torch.testing.assert_close(d.float().cpu(), uu.grad, rtol=tol, atol=tol)
    dot = torch.zeros(1, device="cuda")
    o = torch.empty(m, n, device="cuda", dtype=dtype)
    pf.gemm_nt(A2, B2, dot_with=w.cuda().to(dtype), dot_out=dot, scale_dev=s.cuda(), scale=2.0, act=pf.SA_ACT_MUL_PRE, pre=d, out_act=o)
    v = a2 @ b2.t()
    want = v * 0.74 * d.float().cpu()
    torch.testing.assert_close(o.float().cpu(), want, rtol=tol, atol=tol * float(want.abs().max()))
    assert abs(float(dot) - float((v * w).sum())) <= 1e-3 * float((v * w).abs().sum()) ** 0.5 * 30 + 1e-3


def test_tcgen05_local_attention_rescale_path():
    """the forward kernel keeps O in TMEM and raises its reference maximum only when a tile's maximum exceeds it by 2^8:
    keys whose magnitude grows along the sequence force that rescale (tcgen05.ld / st of the O lanes) many times"""
    ops, pf = _mods()
    g = torch.Generator().manual_seed(5)
    B, H, N, W, d = 1, 2, 1000, 420, 64
    q = _bf(torch.randn(B, H, N, d, generator=g))
    ramp = (1.0 + 24.0 * torch.arange(N) / N).view(1, 1, N, 1)
    k = _bf(torch.randn(B, H, N, d, generator=g) * ramp)
    v = _bf(torch.randn(B, H, N, d, generator=g))
    out = po.local_attention(q, k, v, W, "none")
    inner = H * d
    buf = torch.cat([_heads_to_rows(t) for t in (q, k, v)], dim=1).cuda().bfloat16()
    ldsc = pf.local_desc(B, N, H, d, W, 3 * inner, inner, torch.bfloat16)
    O = torch.zeros(B * N, inner, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device="cuda")
    pf.local_attn_fwd(ldsc, buf, 0, inner, 2 * inner, None, O, 0, lse)
    assert ops.last_path() == 2
    _close(_rows_to_heads(O.float().cpu(), B, H), out, 2e-2, "tc local out (rescale path)")
    ops.set_force_simt(True)
    try:
        O2 = torch.zeros_like(O); lse2 = torch.empty_like(lse)
        pf.local_attn_fwd(ldsc, buf, 0, inner, 2 * inner, None, O2, 0, lse2)
    finally:
        ops.set_force_simt(False)
    _close(lse, lse2, 1e-3, "lse tc vs simt (rescale path)")


@pytest.mark.parametrize("m,n,k", [(1000, 2049, 512), (333, 77, 64), (4100, 1001, 128)])
def test_tcgen05_gemm_nt_unaligned_fp32_rows(m, n, k):
    """fp32 output whose rows start at odd 4-byte offsets (the logits, n = 2049): the epilogue transposes 32 x 32 chunks
    through its staging blocks and stores row segments; same numbers as the per-thread epilogue (SA_GEMM_ROW_OUT=0)."""
    import os
    ops, pf = _mods()
    g = torch.Generator().manual_seed(m + n)
    a, b = _bf(torch.randn(m, k, generator=g)), _bf(torch.randn(n, k, generator=g) * 0.1)
    bias = torch.randn(n, generator=g)
    A, Bm = a.cuda().bfloat16(), b.cuda().bfloat16()
    out = torch.full((m, n), 7.0, device="cuda")
    pf.gemm_nt(A, Bm, bias=bias.cuda(), out_f32=out)
    assert ops.last_path() == 2
    want = a @ b.t() + bias
    _close(out, want, 2e-5 * max(1.0, k / 256), "row-segment epilogue")
    os.environ["SA_GEMM_ROW_OUT"] = "0"
    try:
        out2 = torch.empty_like(out)
        pf.gemm_nt(A, Bm, bias=bias.cuda(), out_f32=out2)
    finally:
        del os.environ["SA_GEMM_ROW_OUT"]
    assert torch.equal(out, out2)


def test_weight_prep_multi_tensor():
    """sa_weight_prep: bf16 copies and transposes of many fp32 matrices in one launch, several sources into one destination"""
    import ctypes as C
    from synthanatomy_b200 import _lib
    ops, pf = _mods()
    g = torch.Generator().manual_seed(9)
    shapes = [(70, 33), (128, 512), (5, 1000), (257, 64)] * 15          # 60 items: more than one launch
    srcs = [torch.randn(r, c, generator=g).cuda() for r, c in shapes]
    dsts = [torch.zeros(r, c, device="cuda", dtype=torch.bfloat16) for r, c in shapes]
    dts = [torch.zeros(c, r, device="cuda", dtype=torch.bfloat16) for r, c in shapes]
    # two more sources stacked into one [6 x 40] destination and its [40 x 6] transpose
    a, b = torch.randn(2, 40, generator=g).cuda(), torch.randn(4, 40, generator=g).cuda()
    cat, cat_t = torch.zeros(6, 40, device="cuda", dtype=torch.bfloat16), torch.zeros(40, 6, device="cuda", dtype=torch.bfloat16)
    items = (_lib.WPrepItem * (len(shapes) + 2))()
    for it, s, d, t in zip(items, srcs, dsts, dts):
        it.src, it.dst, it.dst_t = s.data_ptr(), d.data_ptr(), t.data_ptr()
        it.rows, it.cols, it.dst_ld, it.dst_t_ld = s.shape[0], s.shape[1], s.shape[1], s.shape[0]
    for it, s, r0 in ((items[len(shapes)], a, 0), (items[len(shapes) + 1], b, 2)):
        it.src, it.dst, it.dst_t = s.data_ptr(), cat[r0:].data_ptr(), cat_t[:, r0:].data_ptr()
        it.rows, it.cols, it.dst_ld, it.dst_t_ld = s.shape[0], 40, 40, 6
    _lib.check(ops.lib().sa_weight_prep(items, len(items), torch.cuda.current_stream().cuda_stream), "sa_weight_prep")
    for s, d, t in zip(srcs, dsts, dts):
        assert torch.equal(d, s.bfloat16()) and torch.equal(t, s.bfloat16().t())
    want = torch.cat((a, b)).bfloat16()
    assert torch.equal(cat, want) and torch.equal(cat_t, want.t())
