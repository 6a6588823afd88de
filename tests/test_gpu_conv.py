"""GPU parity of the gather-GEMM conv primitive (through the C ABI) against the CPU oracle (torch fp32 functional
convs pinned by oracle/vqvae_oracle.py:conv3d_naive).  fp32 path: 1e-4 (north_star tolerance).  bf16 tcgen05 path:
compared with the oracle evaluated on bf16-rounded operands (fp32 accumulate) -- the kernel's exact arithmetic
model -- to bf16 output rounding (2^-8 relative), and with the CUDA-core kernel on identical bf16 inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from synthanatomy_b200 import ops
    return ops


def _to_ndhwc(x, dtype):
    return x.permute(0, 2, 3, 4, 1).contiguous().to(dtype).cuda()


def _from_ndhwc(y):
    return y.float().cpu().permute(0, 4, 1, 2, 3).contiguous()


def _ref_fwd(kind, x, w, b, s, p):
    if kind == "conv":
        return F.conv3d(x, w, b, stride=s, padding=p)
    return F.conv_transpose3d(x, w, b, stride=s, padding=p)


CASES_F32 = [
    # kind, cin, cout, k, s, p, (B, D, H, W)
    ("conv", 1, 32, 4, 2, 1, (2, 8, 12, 8)),
    ("conv", 16, 24, 3, 1, 1, (1, 5, 6, 7)),
    ("conv", 24, 16, 1, 1, 0, (2, 4, 5, 3)),
    ("conv", 8, 8, 4, 2, 1, (1, 6, 8, 10)),
    ("deconv", 16, 8, 4, 2, 1, (2, 3, 5, 4)),
    ("deconv", 32, 1, 4, 2, 1, (1, 4, 6, 5)),
]


@pytest.mark.parametrize("kind,cin,cout,k,s,p,shape", CASES_F32)
def test_fp32_fwd_dgrad_wgrad(kind, cin, cout, k, s, p, shape):
    ops = _ops()
    g = torch.Generator().manual_seed(hash((kind, cin, cout, k)) % 1000)
    B, D, H, W = shape
    x = torch.randn(B, cin, D, H, W, generator=g, requires_grad=True)
    wshape = (cout, cin, k, k, k) if kind == "conv" else (cin, cout, k, k, k)
    w = (torch.randn(wshape, generator=g) * 0.1).requires_grad_(True)
    b = torch.randn(cout, generator=g).requires_grad_(True)
    y = _ref_fwd(kind, x, w, b, s, p)
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy)

    spec = ops.ConvSpec(kind, cin, cout, k, s, p)
    xd = _to_ndhwc(x.detach(), torch.float32)
    wd = w.detach().cuda()
    wp = ops.pack_weight(wd, transpose=(kind == "deconv"), dtype=torch.float32)
    yd = ops.conv_forward(spec, xd, wp, b.detach().cuda(), None, False)
    assert ops.last_path() == 1
    torch.testing.assert_close(_from_ndhwc(yd), y.detach(), rtol=1e-4, atol=1e-4)

    gyd = _to_ndhwc(gy, torch.float32)
    wp_t = ops.pack_weight(wd, transpose=(kind == "conv"), dtype=torch.float32)
    dx = ops.conv_dgrad(spec, gyd, wp_t, (D, H, W))
    torch.testing.assert_close(_from_ndhwc(dx), x.grad, rtol=1e-4, atol=1e-4)
    dw = ops.conv_wgrad(spec, xd, gyd, wd)
    torch.testing.assert_close(dw.cpu(), w.grad, rtol=1e-4, atol=2e-4)
    db = ops.bias_grad(gyd)
    torch.testing.assert_close(db.cpu(), b.grad, rtol=1e-4, atol=2e-4)


def test_fp32_epilogue_addend_relu_mask():
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 8, 4, 4, 4, generator=g)
    w = torch.randn(8, 8, 3, 3, 3, generator=g) * 0.1
    b = torch.randn(8, generator=g)
    add = torch.randn(1, 8, 4, 4, 4, generator=g)
    msk = torch.randn(1, 8, 4, 4, 4, generator=g)
    spec = ops.ConvSpec("conv", 8, 8, 3, 1, 1)
    wp = ops.pack_weight(w.cuda(), False, torch.float32)
    y = ops.conv_forward(spec, _to_ndhwc(x, torch.float32), wp, b.cuda(), _to_ndhwc(add, torch.float32), True)
    ref = F.relu(F.conv3d(x, w, b, padding=1) + add)
    torch.testing.assert_close(_from_ndhwc(y), ref, rtol=1e-4, atol=1e-4)
    # dgrad-style epilogue: (acc + addend) * (mask > 0)
    wp_t = ops.pack_weight(w.cuda(), True, torch.float32)
    dx = ops.conv_dgrad(spec, _to_ndhwc(x, torch.float32), wp_t, (4, 4, 4), _to_ndhwc(add, torch.float32),
                        _to_ndhwc(msk, torch.float32))
    ref = (F.conv_transpose3d(x, w, None, padding=1) + add) * (msk > 0)
    torch.testing.assert_close(_from_ndhwc(dx), ref, rtol=1e-4, atol=1e-4)


CASES_TC = [
    # kind, cin, cout, k, s, p, (B, D, H, W)
    ("conv", 64, 64, 1, 1, 0, (1, 4, 8, 8)),        # single tap, single k-chunk
    ("conv", 128, 128, 1, 1, 0, (2, 4, 8, 16)),
    ("conv", 128, 128, 3, 1, 1, (1, 8, 8, 16)),     # the dominant layer type
    ("conv", 128, 128, 3, 1, 1, (2, 5, 7, 10)),     # ragged: tiles overhang every edge
    ("conv", 256, 256, 3, 1, 1, (1, 4, 6, 10)),     # N = 256
    ("conv", 256, 32, 3, 1, 1, (2, 5, 7, 5)),       # pre-quant projection (N = 32; its dgrad gathers 32 channels)
    ("conv", 32, 256, 3, 1, 1, (1, 5, 7, 5)),       # post-quant projection: ONE partial 64-channel TMA box (zero-filled)
    ("conv", 128, 128, 4, 2, 1, (1, 8, 16, 16)),    # strided: 8 parity views
    ("conv", 128, 256, 4, 2, 1, (2, 8, 12, 20)),
    ("deconv", 256, 128, 4, 2, 1, (1, 4, 6, 10)),   # 8 output-parity phases
    ("deconv", 128, 128, 4, 2, 1, (2, 4, 8, 8)),
]


def _bf(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("kind,cin,cout,k,s,p,shape", CASES_TC)
def test_tcgen05_fwd_dgrad_wgrad(kind, cin, cout, k, s, p, shape):
    ops = _ops()
    g = torch.Generator().manual_seed(11)
    B, D, H, W = shape
    x = _bf(torch.randn(B, cin, D, H, W, generator=g)).requires_grad_(True)
    wshape = (cout, cin, k, k, k) if kind == "conv" else (cin, cout, k, k, k)
    w = _bf(torch.randn(wshape, generator=g) * 0.05).requires_grad_(True)
    b = torch.randn(cout, generator=g)
    y = _ref_fwd(kind, x, w, b, s, p)
    gy = _bf(torch.randn(y.shape, generator=g))
    y.backward(gy)
    scale = float(y.detach().abs().max())

    spec = ops.ConvSpec(kind, cin, cout, k, s, p)
    xd = _to_ndhwc(x.detach(), torch.bfloat16)
    wd = w.detach().cuda()
    wp = ops.pack_weight(wd, transpose=(kind == "deconv"), dtype=torch.bfloat16)
    yd = ops.conv_forward(spec, xd, wp, b.cuda(), None, False)
    assert ops.last_path() == 2, "tcgen05 kernel was not selected"
    # bf16 output rounding: 2^-8 relative + a little accumulation-order slack
    torch.testing.assert_close(_from_ndhwc(yd), y.detach(), rtol=2 ** -7, atol=2e-3 * scale)
    ops.set_force_simt(True)
    try:
        ys = ops.conv_forward(spec, xd, wp, b.cuda(), None, False)
        assert ops.last_path() == 1
    finally:
        ops.set_force_simt(False)
    # same inputs, same fp32 accumulation: equal up to one bf16 ulp from summation order
    torch.testing.assert_close(yd.float(), ys.float(), rtol=2 ** -7, atol=1e-3 * scale)

    gyd = _to_ndhwc(gy, torch.bfloat16)
    wp_t = ops.pack_weight(wd, transpose=(kind == "conv"), dtype=torch.bfloat16)
    dx = ops.conv_dgrad(spec, gyd, wp_t, (D, H, W))
    # the dgrad primitive gathers over dy (cout channels): whole 64-channel boxes, or one partial box of 16 / 32 / 48
    assert ops.last_path() == (2 if (cout % 64 == 0 or (cout < 64 and cout % 16 == 0)) else 1)
    sx = float(x.grad.abs().max())
    torch.testing.assert_close(_from_ndhwc(dx), x.grad, rtol=2 ** -7, atol=2e-3 * sx)

    dw = ops.conv_wgrad(spec, xd, gyd, wd)
    if cin in (64, 128, 256) and cout % 128 == 0 and (kind == "conv" or cin % 128 == 0):
        pass  # path asserted below where supported
    sw = float(w.grad.abs().max())
    torch.testing.assert_close(dw.cpu(), w.grad, rtol=1e-3, atol=1e-3 * sw)


@pytest.mark.parametrize("kind,cin,cout,k,s,p,shape", [
    ("conv", 128, 128, 3, 1, 1, (2, 6, 8, 16)),
    ("conv", 128, 128, 1, 1, 0, (1, 4, 8, 16)),
    ("conv", 128, 128, 4, 2, 1, (1, 8, 8, 16)),
    ("deconv", 128, 128, 4, 2, 1, (1, 4, 4, 8)),
    ("conv", 256, 256, 3, 1, 1, (1, 4, 4, 8)),
])
def test_tcgen05_wgrad_selected(kind, cin, cout, k, s, p, shape):
    ops = _ops()
    g = torch.Generator().manual_seed(13)
    B, D, H, W = shape
    x = _bf(torch.randn(B, cin, D, H, W, generator=g))
    wshape = (cout, cin, k, k, k) if kind == "conv" else (cin, cout, k, k, k)
    w = torch.zeros(wshape, requires_grad=True)
    y = _ref_fwd(kind, x, w, None, s, p)
    gy = _bf(torch.randn(y.shape, generator=g))
    y.backward(gy)
    spec = ops.ConvSpec(kind, cin, cout, k, s, p)
    dw = ops.conv_wgrad(spec, _to_ndhwc(x, torch.bfloat16), _to_ndhwc(gy, torch.bfloat16), w.detach().cuda())
    assert ops.last_path() == 2, "tcgen05 wgrad kernel was not selected"
    sw = float(w.grad.abs().max())
    torch.testing.assert_close(dw.cpu(), w.grad, rtol=1e-3, atol=1e-3 * sw)


def test_tcgen05_epilogue_fusions():
    ops = _ops()
    g = torch.Generator().manual_seed(17)
    x = _bf(torch.randn(1, 128, 4, 8, 8, generator=g))
    w = _bf(torch.randn(128, 128, 3, 3, 3, generator=g) * 0.05)
    b = torch.randn(128, generator=g)
    add = _bf(torch.randn(1, 128, 4, 8, 8, generator=g))
    msk = _bf(torch.randn(1, 128, 4, 8, 8, generator=g))
    spec = ops.ConvSpec("conv", 128, 128, 3, 1, 1)
    wp = ops.pack_weight(w.cuda(), False, torch.bfloat16)
    y = ops.conv_forward(spec, _to_ndhwc(x, torch.bfloat16), wp, b.cuda(), _to_ndhwc(add, torch.bfloat16), True)
    assert ops.last_path() == 2
    ref = F.relu(F.conv3d(x, w, b, padding=1) + add)
    torch.testing.assert_close(_from_ndhwc(y), ref, rtol=2 ** -7, atol=2e-3 * float(ref.abs().max()))
    wp_t = ops.pack_weight(w.cuda(), True, torch.bfloat16)
    dx = ops.conv_dgrad(spec, _to_ndhwc(x, torch.bfloat16), wp_t, (4, 8, 8), _to_ndhwc(add, torch.bfloat16),
                        _to_ndhwc(msk, torch.bfloat16))
    assert ops.last_path() == 2
    ref = (F.conv_transpose3d(x, w, None, padding=1) + add) * (msk > 0)
    torch.testing.assert_close(_from_ndhwc(dx), ref, rtol=2 ** -7, atol=2e-3 * float(ref.abs().max()))


@pytest.mark.parametrize("shape", [(1, 4, 8, 8), (2, 5, 7, 10), (1, 16, 16, 24)])
def test_tcgen05_pointwise_backward_fused(shape):
    """sa_conv1x1_bwd_fused (dh with the ReLU mask of h, dW1, db1 in one pass) against torch autograd on the same
    bf16-rounded tensors, and against the three general entry points it replaces."""
    ops = _ops()
    g_ = torch.Generator().manual_seed(23)
    B, D, H, W = shape
    C = 128
    h = _bf(torch.randn(B, C, D, H, W, generator=g_)).requires_grad_(True)     # pre-ReLU sign pattern matters: use h itself
    w = _bf(torch.randn(C, C, 1, 1, 1, generator=g_) * 0.05).requires_grad_(True)
    b = torch.zeros(C, requires_grad=True)
    gy = _bf(torch.randn(B, C, D, H, W, generator=g_))
    hr = F.relu(h)
    y = F.conv3d(hr, w, b)
    y.backward(gy)
    spec = ops.ConvSpec("conv", C, C, 1, 1, 0)
    gd, hd = _to_ndhwc(gy, torch.bfloat16), _to_ndhwc(hr.detach(), torch.bfloat16)
    wp_t = ops.pack_weight(w.detach().cuda(), True, torch.bfloat16)
    assert ops.conv1x1_bwd_fused_supported(spec, gd)
    dh, dw, db = ops.conv1x1_bwd_fused(spec, gd, hd, wp_t, w.detach().cuda())
    assert ops.last_path() == 2
    torch.testing.assert_close(_from_ndhwc(dh), h.grad, rtol=2 ** -7, atol=2e-3 * float(h.grad.abs().max()))
    torch.testing.assert_close(dw.cpu(), w.grad, rtol=1e-3, atol=1e-3 * float(w.grad.abs().max()))
    torch.testing.assert_close(db.cpu(), b.grad, rtol=1e-3, atol=1e-3 * float(b.grad.abs().max()))
    dh2 = ops.conv_dgrad(spec, gd, wp_t, (D, H, W), None, hd)
    assert torch.equal(dh2, dh)
    # the streaming forward twin: relu(conv1x1(h) + b + x) == the general entry point, bit for bit
    bias = torch.randn(C, generator=g_).cuda()
    xd = _to_ndhwc(_bf(torch.randn(B, C, D, H, W, generator=g_)), torch.bfloat16)
    wp = ops.pack_weight(w.detach().cuda(), False, torch.bfloat16)
    y1 = ops.conv1x1_fwd_fused(spec, hd, wp, bias, xd, True)
    y2 = ops.conv_forward(spec, hd, wp, bias, xd, True)
    assert torch.equal(y1, y2)
