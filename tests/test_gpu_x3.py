"""bf16x3: the tensor-core PARITY mode (csrc/sa_x3.cu) -- fp32 tensors, every product on the bf16 tensor cores as
hi.hi + lo.hi + hi.lo of split operands with fp32 accumulation.  Held to the north_star tolerance (1e-4 against the
fp32 CPU oracle; kernels against float64 at 3e-5 (dense) / 5e-5 (convs, up to 8192-term contractions) of the result's
scale), i.e. the same bar as the CUDA-core fp32 path."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mods():
    from synthanatomy_b200 import ops, pf_ops
    return ops, pf_ops


def _rel(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return float((got - want).abs().max()) / max(float(want.abs().max()), 1e-30)


@pytest.mark.parametrize("m,n,k", [(300, 512, 128), (1000, 256, 512), (257, 80, 64), (130, 2049, 192), (4099, 1024, 250),
                                   (2500, 512, 2049)])
def test_gemm_nt_x3_against_float64(m, n, k):
    ops, pf = _mods()
    g = torch.Generator().manual_seed(m + n + k)
    a, b = torch.randn(m, k, generator=g), torch.randn(n, k, generator=g) * 0.1
    bias, resid, w = torch.randn(n, generator=g), torch.randn(m, n, generator=g), torch.randn(m, n, generator=g)
    s = torch.tensor([0.37])
    A, B = a.cuda(), b.cuda()
    v = a.double() @ b.double().t()
    with ops.x3_mode(True):
        out = torch.empty(m, n, device="cuda")
        pf.gemm_nt(A, B, out_f32=out)
        assert ops.last_path() == 2, "the tcgen05 GEMM was not selected"
        assert _rel(out, v) <= 3e-5, _rel(out, v)
        # the plain bf16 tensor-core product of the same operands is 100x further away: the split is doing the work
        out16 = torch.empty(m, n, device="cuda")
        pf.gemm_nt(A.bfloat16(), B.bfloat16(), out_f32=out16)
        assert _rel(out16, v) > 20 * _rel(out, v)
        # FFN-1 epilogue: bias + GELU, fp32 pre-activation kept
        pre, h = torch.empty(m, n, device="cuda"), torch.empty(m, n, device="cuda")
        pf.gemm_nt(A, B, bias=bias.cuda(), act=pf.SA_ACT_GELU_FWD, pre=pre, out_act=h)
        u = v + bias.double()
        assert _rel(pre, u) <= 3e-5 and _rel(h, F.gelu(u)) <= 3e-5
        # ReZero residual epilogue (in place)
        r = resid.cuda().clone()
        pf.gemm_nt(A, B, bias=bias.cuda(), scale_dev=s.cuda(), resid=r, out_f32=r)
        assert _rel(r, resid.double() + 0.37 * u) <= 3e-5
        # backward-style epilogue: dot with a tensor, scale, GELU'
        dot = torch.zeros(1, device="cuda")
        o3 = torch.empty(m, n, device="cuda")
        pf.gemm_nt(A, B, dot_with=w.cuda(), dot_out=dot, scale_dev=s.cuda(), scale=2.0, act=pf.SA_ACT_GELU_BWD, pre=pre, out_act=o3)
        uu = u.clone().requires_grad_(True)
        F.gelu(uu).sum().backward()
        assert abs(float(dot) - float((v * w.double()).sum())) <= 1e-5 * float((v * w.double()).abs().sum())
        assert _rel(o3, v * 0.74 * uu.grad) <= 5e-5


def test_gemm_tn_x3_against_float64_with_column_slices():
    ops, pf = _mods()
    g = torch.Generator().manual_seed(3)
    big_a, big_b = torch.randn(5000, 200, generator=g), torch.randn(5000, 300, generator=g)
    a, b = big_a[:, 30:100], big_b[:, 64:194]
    A, B = big_a.cuda()[:, 30:100], big_b.cuda()[:, 64:194]
    s = torch.tensor([-0.5], device="cuda")
    want = a.double().t() @ b.double()
    with ops.x3_mode(True):
        d = torch.empty(70, 130, device="cuda")
        pf.gemm_tn(A, B, d, scale_dev=s, scale=2.0)
        assert ops.last_path() == 2
        assert _rel(d, -want) <= 3e-5
        pf.gemm_tn(A, B, d, accumulate=True)
        assert float(d.abs().max()) <= 1e-4 * float(want.abs().max())


CONV_CASES = [
    # kind, cin, cout, k, s, p, (B, D, H, W)
    ("conv", 128, 128, 3, 1, 1, (2, 9, 6, 10)),
    ("conv", 128, 128, 1, 1, 0, (1, 5, 7, 9)),
    ("conv", 128, 128, 4, 2, 1, (2, 8, 12, 8)),
    ("conv", 256, 32, 3, 1, 1, (1, 5, 7, 5)),
    ("conv", 32, 128, 3, 1, 1, (1, 6, 5, 7)),
    ("deconv", 128, 128, 4, 2, 1, (1, 4, 6, 5)),
    ("deconv", 256, 128, 4, 2, 1, (1, 3, 4, 5)),
    ("conv", 256, 256, 3, 1, 1, (1, 5, 7, 5)),           # 256 output channels: two launches over the channel halves
    ("conv", 128, 256, 4, 2, 1, (1, 8, 4, 12)),
]


@pytest.mark.parametrize("kind,cin,cout,k,s,p,shape", CONV_CASES)
def test_conv_x3_fwd_dgrad_wgrad_against_float64(kind, cin, cout, k, s, p, shape):
    ops, _ = _mods()
    g = torch.Generator().manual_seed(cin + cout + k)
    B, D, H, W = shape
    x = torch.randn(B, cin, D, H, W, generator=g).double().requires_grad_(True)
    wshape = (cout, cin, k, k, k) if kind == "conv" else (cin, cout, k, k, k)
    w = (torch.randn(wshape, generator=g) * 0.05).double().requires_grad_(True)
    b = torch.randn(cout, generator=g).double()
    y = F.conv3d(x, w, b, stride=s, padding=p) if kind == "conv" else F.conv_transpose3d(x, w, b, stride=s, padding=p)
    gy = torch.randn(y.shape, generator=g).double()
    y.backward(gy)

    def nd(t):
        return t.detach().float().permute(0, 2, 3, 4, 1).contiguous().cuda()

    def nc(t):
        return t.float().cpu().permute(0, 4, 1, 2, 3)

    spec = ops.ConvSpec(kind, cin, cout, k, s, p)
    wd = w.detach().float().cuda()
    f32 = torch.float32
    square = cin == cout == 128       # the layers that carry 95 % of the FLOPs: all three kernels must be the tcgen05 ones
    with ops.x3_mode(True):           # (shapes the tensor-core kernels do not take run on the CUDA-core fp32 kernels)
        yd = ops.conv_forward(spec, nd(x), ops.pack_weight(wd, kind == "deconv", f32), b.float().cuda(), None, False)
        assert ops.last_path() == 2, "forward did not run on the tcgen05 kernel"
        assert _rel(nc(yd), y) <= 5e-5, ("fwd", _rel(nc(yd), y))
        dx = ops.conv_dgrad(spec, nd(gy), ops.pack_weight(wd, kind == "conv", f32), (D, H, W))
        assert ops.last_path() == 2 or not square, "dgrad did not run on the tcgen05 kernel"
        assert _rel(nc(dx), x.grad) <= 5e-5, ("dgrad", _rel(nc(dx), x.grad))
        dw = ops.conv_wgrad(spec, nd(x), nd(gy), wd)
        assert ops.last_path() == 2 or not square, "wgrad did not run on the tcgen05 kernel"
        assert _rel(dw, w.grad) <= 5e-5, ("wgrad", _rel(dw, w.grad))
    # epilogue: bias + addend + ReLU + mask on fp32 tensors
    if kind == "conv" and k == 3 and cin == cout:
        add = torch.randn(y.shape, generator=g)
        msk = torch.randn(y.shape, generator=g)
        with ops.x3_mode(True):
            yd = ops.conv_forward(spec, nd(x), ops.pack_weight(wd, False, f32), b.float().cuda(), nd(add), True, nd(msk))
        want = torch.relu(y.detach() + add.double()) * (msk > 0)
        assert _rel(nc(yd), want) <= 3e-5


def test_vqvae_bf16x3_meets_the_fp32_tolerance():
    """2-level, 256 channels (128-channel level-1 layers on the split tensor-core path, 256-output-channel layers and the
    1-channel ends on the CUDA-core fp32 kernels): reconstruction, loss, code indices and every gradient against the
    fp32 oracle at the north_star tolerance"""
    from oracle import vqvae_oracle as vo
    from synthanatomy_b200 import ops
    from synthanatomy_b200.networks.vqvae import B200VQVAE
    kw = dict(n_levels=2, downsample_parameters=((4, 2, 1, 1),) * 2, upsample_parameters=((4, 2, 1, 0, 1),) * 2,
              n_embed=64, embed_dim=32, n_channels=256, n_res_channels=256, n_res_layers=2, vq_decay=0.5,
              commitment_cost=0.25)
    torch.manual_seed(1)
    net = B200VQVAE(**kw, compute_dtype=ops.BF16X3)
    with torch.no_grad():
        net.quantizer[0].impl.embedding.weight.mul_(0.05)
        net.quantizer[0].impl.embed_avg.copy_(net.quantizer[0].impl.embedding.weight)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    x = torch.rand(2, 1, 16, 24, 32)
    loss_ref, grads_ref, out_ref = vo.train_step_grads(sd, vo.VQVAEConfig(**kw), x)
    net = net.cuda()
    with torch.no_grad():       # before the training step below moves the codebook (EMA)
        assert torch.equal(net.eval().index_quantize(x.cuda())[0].cpu(), out_ref["indices"])
    net.train()
    ops.reset_launch_count()
    out = net(x.cuda())
    loss = F.mse_loss(out["reconstruction"][0], x.cuda()) + out["quantization_losses"][0]
    loss.backward()
    assert float((out["reconstruction"][0].detach().cpu() - out_ref["reconstruction"][0]).abs().max()) <= 1e-4
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * max(1.0, abs(float(loss_ref)))
    for k, p in net.named_parameters():
        if p.requires_grad:
            err = float((p.grad.cpu() - grads_ref[k]).abs().max())
            assert err <= 2e-4 * max(float(grads_ref[k].abs().max()), 1e-3), f"{k}: {err:.3e}"


def test_performer_bf16x3_meets_the_fp32_tolerance():
    from oracle import performer_oracle as po
    from synthanatomy_b200 import ops
    from synthanatomy_b200.losses import CELoss
    from tests.test_gpu_performer import CASES, _build, _close
    kw, grid = CASES["readme_slice"]
    cfg, sd, net, seqs, x_in, y = _build(kw, grid, 11, compute_dtype=ops.BF16X3)
    loss_ref, grads_ref, logits_ref = po.train_step_grads(sd, cfg, x_in, y, seqs)
    net = net.cuda().train()
    logits = net(x_in.cuda())
    assert ops.last_path() == 2, "the logits GEMM did not run on the tcgen05 kernel"
    _close(logits, logits_ref, 1e-4, "bf16x3 logits")
    loss = CELoss()(logits.transpose(1, 2), y.cuda())
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * max(1.0, abs(float(loss_ref)))
    loss.backward()
    named = dict(net.named_parameters())
    for k, gref in grads_ref.items():
        err = float((named[k].grad.cpu() - gref).abs().max()) / max(float(gref.abs().max()), 1e-3)
        assert err <= 2e-4, f"grad {k}: {err:.3e}"
