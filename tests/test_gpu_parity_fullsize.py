"""Oracle parity of the BENCHMARKED path (bf16 operands on tcgen05, fp32 accumulation) at BASELINE.json's full shapes.

The CPU oracle cannot run config 2 at batch 8 in seconds, but it can at batch 1 (the bench's own `cpu_baseline` leg
runs one 160 x 224 x 160 volume in ~9 s).  So:

  * every conv of the 4-level / 256-channel VQ-VAE at its true 160 x 224 x 160 shape, TEACHER-FORCED: the fp32 oracle
    (oracle/vqvae_oracle.py, pinned to the reference) walks the network once; each layer of the product then runs on the
    oracle's own input of that layer and is held to the oracle's output of that layer evaluated on the SAME bf16-rounded
    operands (fp32 accumulation) -- the kernel's exact arithmetic model -- to 2 bf16 ulp per element;
  * the data / weight / bias gradients of the layers that dominate the step (level-1 ResidualLayer, the 4/2/1 strided
    conv and transposed conv, both 1-channel ends) against autograd of the oracle's functional layer on the same
    operands;
  * the whole model, forward + backward at 160 x 224 x 160 (batch 1) against ``vo.train_step_grads`` in fp32;
  * the Performer at N = 14 000 (grid 20 x 28 x 25, dim 512, 16 heads of which 8 local, window 420), depth 2, against
    ``po.train_step_grads``.

Tolerances are stated as bf16 ulps (2^-8 relative) of the quantity compared, not as an end-to-end constant.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ULP = 2.0 ** -8        # bf16: 8 significand bits => relative rounding error <= 2^-9, spacing 2^-8

KW = dict(n_levels=4, downsample_parameters=((4, 2, 1, 1),) * 4, upsample_parameters=((4, 2, 1, 0, 1),) * 4,
          n_embed=2048, embed_dim=32, n_channels=256, n_res_channels=256, n_res_layers=3, vq_decay=0.5,
          commitment_cost=0.25)
VOL = (160, 224, 160)


def _bf(t):
    return t.to(torch.bfloat16).float()


def _ndhwc(x, dtype=torch.bfloat16):
    return x.permute(0, 2, 3, 4, 1).contiguous().to(dtype).cuda()


def _ncdhw(y):
    return y.float().cpu().permute(0, 4, 1, 2, 3).contiguous()


def _assert_ulps(got, want, ulps, what):
    """|got - want| <= ulps * 2^-8 * max(|want|, rms(want)) elementwise"""
    got, want = got.float(), want.float()
    rms = float(want.pow(2).mean().sqrt())
    bound = ulps * ULP * torch.maximum(want.abs(), torch.full_like(want, rms))
    excess = (got - want).abs() - bound
    worst = float(excess.max())
    assert worst <= 0, (f"{what}: exceeds {ulps} bf16 ulp by {worst:.3e} (rms {rms:.3e}, max |want| "
                        f"{float(want.abs().max()):.3e})")


def _model():
    from oracle import vqvae_oracle as vo
    from synthanatomy_b200.networks.vqvae import B200VQVAE
    torch.manual_seed(4)
    net = B200VQVAE(**KW, compute_dtype=torch.bfloat16)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    return vo, vo.VQVAEConfig(**KW), net, sd


def _layer_walk(vo, cfg, sd, x):
    """fp32 oracle walk: one entry (stack, kind, params-prefix, meta, input, output) per conv / ResidualLayer, i.e. per op
    of the product's stack programme; kind is 'conv' | 'deconv' | 'res' (ONE residual layer, prefix '<group>.<j>')"""
    out = []

    def run(stack, prog, prefix, h):
        for kind, i, m in prog:
            p = f"{prefix}.{i}"
            if kind == "res":
                for j in range(m["n"]):
                    y = _oracle_op(sd, f"{p}.{j}", "res", m, h)
                    out.append((stack, "res", f"{p}.{j}", m, h, y))
                    h = y
            else:
                y = _oracle_op(sd, p, kind, m, h)
                out.append((stack, kind, p, m, h, y))
                h = y
        return h

    with torch.no_grad():
        z = run("enc", vo.encoder_program(cfg), "encoder.0", x)
        q_st, _, _, _ = vo.quantize(sd, cfg, z, True)
        run("dec", vo.decoder_program(cfg), "decoder.0", q_st)
    return out


def _oracle_op(sd, p, kind, m, x, rounded=False):
    """one conv / ResidualLayer of the oracle (baseline.py:150-160, 218-228, 242-244, 258, 283-297); rounded=True evaluates
    it in the bf16 kernels' arithmetic model: bf16 operands, fp32 accumulation, bf16 activation between the two convs"""
    r = _bf if rounded else (lambda t: t)
    if kind == "conv":
        y = F.conv3d(r(x), r(sd[f"{p}.weight"]), sd[f"{p}.bias"], stride=m["s"], padding=m["p"])
        return F.relu(y) if m["relu"] else y
    if kind == "deconv":
        y = F.conv_transpose3d(r(x), r(sd[f"{p}.weight"]), sd[f"{p}.bias"], stride=m["s"], padding=m["p"])
        return F.relu(y) if m["relu"] else y
    h = r(x)
    t = F.relu(F.conv3d(h, r(sd[f"{p}.0.weight"]), sd[f"{p}.0.bias"], padding=1))
    t = F.conv3d(r(t), r(sd[f"{p}.3.weight"]), sd[f"{p}.3.bias"])
    return F.relu(h + t)


def test_vqvae_config2_every_layer_teacher_forced_against_oracle():
    """all 58 convs (34 programme ops: 10 convs + 24 ResidualLayers) of BASELINE config 2 at 1 x 160 x 224 x 160, bf16
    tcgen05 path, each on the oracle's own layer input.  Single convs: 2 ulp.  A ResidualLayer is 3x3x3 -> ReLU -> [bf16]
    -> 1x1x1 -> +x -> ReLU: an element of the intermediate h that sits on a rounding boundary may land one bf16 step
    apart between the two fp32 accumulation orders, which the 1x1x1 conv then spreads over the output channels, so a
    layer is held to 4 ulp."""
    from synthanatomy_b200 import ops
    from synthanatomy_b200.networks.vqvae import b200
    vo, cfg, net, sd = _model()
    x = torch.rand(1, 1, *VOL, generator=torch.Generator().manual_seed(7))
    walk = _layer_walk(vo, cfg, sd, x)
    net = net.cuda()
    progs = {"enc": net._enc_ops, "dec": net._dec_ops}
    cursor = {"enc": 0, "dec": 0}
    tc_ops = 0
    report = []
    for stack, kind, p, m, xin, _y32 in walk:
        sub = progs[stack][cursor[stack]: cursor[stack] + 1]
        cursor[stack] += 1
        assert isinstance(sub[0], b200._ResOp) == (kind == "res"), p
        params = [t.detach().float().contiguous() for t in sub[0].params()]
        with torch.no_grad():
            want = _bf(_oracle_op(sd, p, kind, m, xin, rounded=True))
            got, _ = b200._stack_forward(sub, _ndhwc(xin), params, save=False)
        tc_ops += int(ops.last_path() == 2)
        ulps = 4 if kind == "res" else 2
        if kind == "deconv" and m["cout"] == 1:
            # last ConvTranspose3d 128 -> 1: the per-tap products r[pos][tap] are a bf16 tensor before the col2im gather
            # sums 8 of them (one more rounding, of terms that may be larger than their sum)
            ulps = 16
        try:
            _assert_ulps(_ncdhw(got), want, ulps, f"{p} ({kind})")
        except AssertionError as e:
            report.append(str(e))
    assert not report, "\n".join(report)
    assert cursor["enc"] == len(net._enc_ops) and cursor["dec"] == len(net._dec_ops)
    assert tc_ops >= 28, "the tcgen05 kernels were not the ones exercised"


def _grad_case(sd, p, kind, m, xin, seed):
    """autograd of the oracle's functional layer on bf16-rounded operands; returns tensors for the GPU side"""
    g = torch.Generator().manual_seed(seed)
    xr = _bf(xin).requires_grad_(True)
    leaves = {}

    def leaf(k, rounded=True):
        t = sd[k]
        leaves[k] = (_bf(t) if rounded else t.clone()).requires_grad_(True)
        return leaves[k]

    if kind == "res":
        h = F.relu(F.conv3d(xr, leaf(f"{p}.0.weight"), leaf(f"{p}.0.bias", False), padding=1))
        hb = h + (_bf(h.detach()) - h.detach())                 # bf16 activation between the two convs (straight-through)
        pre = xr + F.conv3d(hb, leaf(f"{p}.3.weight"), leaf(f"{p}.3.bias", False))
        y = F.relu(pre)
        aux = (hb.detach(), y.detach())
    elif kind == "conv":
        y = F.conv3d(xr, leaf(f"{p}.weight"), leaf(f"{p}.bias", False), stride=m["s"], padding=m["p"])
        aux = ()
    else:
        y = F.conv_transpose3d(xr, leaf(f"{p}.weight"), leaf(f"{p}.bias", False), stride=m["s"], padding=m["p"])
        aux = ()
    gy = _bf(torch.randn(y.shape, generator=g))
    if kind == "res":
        gy = gy * (y.detach() > 0)                              # the product's backward receives g already ReLU-masked
        pre.backward(gy)
    else:
        y.backward(gy)
    return xr, gy, leaves, aux


def _close_rel(got, want, tol, what):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    err = float((got - want).abs().max())
    scale = float(want.abs().max())
    assert err <= tol * scale, f"{what}: max abs err {err:.3e} > {tol:.1e} * max |want| {scale:.3e}"


def test_vqvae_config2_dominant_layer_gradients_against_oracle():
    """dgrad / wgrad / bias-grad kernels at the true level-1 shapes (1 x 80 x 112 x 80 x 128) and at the 1-channel ends,
    bf16 operands.  dx is a bf16 tensor (2 ulp); dW / db are fp32 sums over 7e5 - 5.7e6 positions (1e-3 of max |dW|:
    fp32 accumulation order on both sides)."""
    from synthanatomy_b200 import ops
    from synthanatomy_b200.ops import ConvSpec
    vo, cfg, net, sd = _model()
    x = torch.rand(1, 1, *VOL, generator=torch.Generator().manual_seed(8))
    walk = _layer_walk(vo, cfg, sd, x)
    pick = {"encoder.0.0": None, "encoder.0.2.0": None, "encoder.0.3": None, "decoder.0.8": None, "decoder.0.10.2": None,
            "decoder.0.11": None}
    for stack, kind, p, m, xin, y in walk:
        if p in pick:
            pick[p] = (kind, m, xin)
    assert all(v is not None for v in pick.values()), [k for k, v in pick.items() if v is None]
    bf = torch.bfloat16
    for seed, (p, (kind, m, xin)) in enumerate(pick.items()):
        xr, gy, leaves, aux = _grad_case(sd, p, kind, m, xin, 100 + seed)
        gd = _ndhwc(gy)
        xd = _ndhwc(xr.detach())
        in_dhw = tuple(xin.shape[2:])
        if kind == "res":
            hb, _y = aux
            w3, w1 = leaves[f"{p}.0.weight"], leaves[f"{p}.3.weight"]
            c = m["c"]
            s3, s1 = ConvSpec("conv", c, c, 3, 1, 1), ConvSpec("conv", c, c, 1, 1, 0)
            hd = _ndhwc(hb)
            wp1_t = ops.pack_weight(w1.detach().cuda(), True, bf)
            assert ops.conv1x1_bwd_fused_supported(s1, gd)
            dh, dw1, db1 = ops.conv1x1_bwd_fused(s1, gd, hd, wp1_t, w1.detach().cuda())
            dw3 = ops.conv_wgrad(s3, xd, dh, w3.detach().cuda())
            db3 = ops.bias_grad(dh)
            wp3_t = ops.pack_weight(w3.detach().cuda(), True, bf)
            dx = ops.conv_dgrad(s3, dh, wp3_t, in_dhw, gd, None)              # (+ g): the residual branch
            assert ops.last_path() == 2
            _close_rel(dw1, w1.grad, 1e-3, f"{p} dW1"); _close_rel(db1, leaves[f"{p}.3.bias"].grad, 1e-3, f"{p} db1")
            # dW3 / db3 / dx see dh rounded to bf16 (one more rounding than the oracle's fp32 chain)
            _close_rel(dw3, w3.grad, ULP, f"{p} dW3"); _close_rel(db3, leaves[f"{p}.0.bias"].grad, ULP, f"{p} db3")
            _assert_ulps(_ncdhw(dx), xr.grad, 6, f"{p} dx")
            continue
        w = leaves[f"{p}.weight"]
        op = [o for o in (net._enc_ops + net._dec_ops) if isinstance(o, type(net._enc_ops[0])) and o.module.weight.shape ==
              w.shape and o.spec.kind == kind][0]
        sp = op.spec
        if op.single_channel_gemm(bf):
            # the product's own path for the 1-channel ends: im2col / col2im + a tcgen05 GEMM over the 64 taps
            from synthanatomy_b200.networks.vqvae import b200
            params = [w.detach().cuda().contiguous(), leaves[f"{p}.bias"].detach().cuda().contiguous()]
            with torch.no_grad():
                _yd, saved = b200._stack_forward([op], xd, params, save=True)
                # _stack_backward refuses stacks that end in a ReLU; these two ops are evaluated without it here
                relu, op.relu = op.relu, False
                try:
                    dx, grads = b200._stack_backward([op], saved, params, gd, False, need_dx=(kind == "deconv"))
                finally:
                    op.relu = relu
            _close_rel(grads[0], w.grad, 1e-3, f"{p} dW"); _close_rel(grads[1], leaves[f"{p}.bias"].grad, 1e-3, f"{p} db")
            if dx is not None:
                _assert_ulps(_ncdhw(dx), xr.grad, 2, f"{p} dx")
            continue
        dw = ops.conv_wgrad(sp, xd, gd, w.detach().cuda())
        db = ops.bias_grad(gd)
        wp_t = ops.pack_weight(w.detach().cuda(), transpose=(kind == "conv"), dtype=bf)
        dx = ops.conv_dgrad(sp, gd, wp_t, in_dhw)
        assert ops.last_path() == 2
        _close_rel(dw, w.grad, 1e-3, f"{p} dW"); _close_rel(db, leaves[f"{p}.bias"].grad, 1e-3, f"{p} db")
        _assert_ulps(_ncdhw(dx), xr.grad, 2, f"{p} dx")


def _conditioned_case():
    """config-2 model + one volume + a codebook on which the argmin is stable under rounding noise (see the test below)"""
    vo, cfg, net, sd = _model()
    x = torch.rand(1, 1, *VOL, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        z = vo.encode(sd, cfg, x)
        lat = z.permute(0, 2, 3, 4, 1).reshape(-1, 32)                      # 1400 latents
        g = torch.Generator().manual_seed(10)
        n_own = lat.shape[0]
        own = lat[torch.randperm(n_own, generator=g)]
        cbw = torch.cat((own + 0.25 * lat.std() * torch.randn(n_own, 32, generator=g),
                         lat.mean(0) + lat.std(0) * torch.randn(2048 - n_own, 32, generator=g)))
        for k in ("quantizer.0.impl.weight", "quantizer.0.impl.embedding.weight", "quantizer.0.impl.embed_avg"):
            sd[k] = cbw.clone()
    return vo, cfg, net, sd, x


def test_vqvae_config2_full_model_step_against_oracle():
    """BASELINE config 2 (4 levels, 256 channels, codebook 2048 x 32) at 1 x 160 x 224 x 160, training mode: forward,
    loss, perplexity, code indices, EMA codebook and every parameter gradient, bf16 tcgen05 path vs the fp32 oracle.
    The encoder is 29 bf16 convs deep, so the latents carry ~sqrt(29) * 2^-9 = 1e-2 relative noise; a latent that close to
    a Voronoi face picks the neighbouring code, after which the decoder sees a different input at that position (a
    discrete event, not a rounding error).  The codebook is built so that such faces are rare (half the codes sit on
    latents of this very volume), the remaining flips are counted, and the bounds are stated for what is left:
    >= 97 % identical indices, reconstruction within 2e-2 relative L2, loss within 2e-2, every gradient cosine >= 0.97."""
    vo, cfg, net, sd, x = _conditioned_case()
    net.load_state_dict(sd)
    loss_ref, grads_ref, out_ref = vo.train_step_grads(sd, cfg, x)
    net = net.cuda().train()
    with torch.no_grad():
        net.eval()
        idx = net.index_quantize(x.cuda())[0].cpu()
        net.train()
    same = float((idx == out_ref["indices"]).float().mean())
    out = net(x.cuda())
    rec = out["reconstruction"][0]
    loss = F.mse_loss(rec.float(), x.cuda()) + out["quantization_losses"][0]
    loss.backward()
    rec_ref = out_ref["reconstruction"][0]
    rel = float((rec.detach().cpu() - rec_ref).norm() / rec_ref.norm())
    cosines = {}
    for k, p in net.named_parameters():
        if p.requires_grad:
            a, b = p.grad.cpu().flatten().double(), grads_ref[k].flatten().double()
            cosines[k] = float((a @ b) / (a.norm() * b.norm() + 1e-30))
    worst = min(cosines, key=cosines.get)
    summary = (f"indices equal {same:.4f}; recon rel L2 {rel:.3e}; loss {float(loss):.6f} vs {float(loss_ref):.6f}; "
               f"min grad cosine {cosines[worst]:.4f} ({worst})")
    print(summary)
    assert same >= 0.97, summary
    assert rel <= 1e-2, summary
    assert abs(float(loss) - float(loss_ref)) <= 1e-2 * abs(float(loss_ref)), summary
    ppl = float(net.get_perplexity()[0])
    ppl_ref = float(vo.perplexity(out_ref["indices"], cfg.n_embed))
    assert abs(ppl - ppl_ref) <= 0.05 * ppl_ref, (ppl, ppl_ref)
    assert cosines[worst] >= 0.99, summary
    # EMA statistics: the codebook rows the oracle updated are the rows this path updated (same argmin up to near-ties)
    n_ref = out_ref["new_state"]["N"]
    n_got = net.quantizer[0].impl.N.cpu()
    agree = float(((n_ref > 0) == (n_got > 0)).float().mean())
    assert agree >= 0.97, f"EMA cluster-usage pattern agreement {agree:.3f}"


def test_vqvae_config2_full_model_step_bf16x3_meets_1e4():
    """the same full-size step in the tensor-core PARITY mode (compute_dtype = BF16X3: fp32 tensors, split-bf16 products
    on tcgen05, fp32 accumulation): held to the north_star tolerance against the fp32 oracle -- reconstruction and loss
    1e-4, identical code indices; parameter gradients to the ReLU-boundary noise floor of this size (see below)."""
    from synthanatomy_b200 import ops
    vo, cfg, net, sd, x = _conditioned_case()
    net.compute_dtype = ops.BF16X3
    net.load_state_dict(sd)
    loss_ref, grads_ref, out_ref = vo.train_step_grads(sd, cfg, x)
    net = net.cuda()
    with torch.no_grad():
        idx = net.eval().index_quantize(x.cuda())[0].cpu()
    assert torch.equal(idx, out_ref["indices"])
    net.train()
    out = net(x.cuda())
    rec = out["reconstruction"][0]
    loss = F.mse_loss(rec, x.cuda()) + out["quantization_losses"][0]
    loss.backward()
    rec_ref = out_ref["reconstruction"][0]
    err = float((rec.detach().cpu() - rec_ref).abs().max())
    assert err <= 1e-4 * max(1.0, float(rec_ref.abs().max())), f"reconstruction max abs err {err:.3e}"
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * max(1.0, abs(float(loss_ref)))
    # Gradients at this size: the forward agrees to ~1e-5, but 1.6e9 activations sit behind ReLUs and a fraction f ~ 1e-5
    # of them lies within that distance of zero, so their masks differ between ANY two fp32 evaluation orders.  Each such
    # element switches one full term of a position sum on or off: a weight gradient (a sign-alternating sum over N
    # positions, |sum| ~ sqrt(N) x term) moves by ~sqrt(f N) x term, i.e. by ~sqrt(f) ~ 3e-3 of its scale (measured run to
    # run: 3e-4 .. 1.5e-3 of max |grad|, on a different tensor each time because the split-K reductions are atomic).  The
    # small-shape tests hold every gradient to 2e-4; here: 5e-3 of max |grad| and 3e-3 in relative L2 per tensor.
    worst = worst_l2 = 0.0
    for k, p in net.named_parameters():
        if p.requires_grad:
            d = (p.grad.cpu() - grads_ref[k]).double()
            e = float(d.abs().max()) / max(float(grads_ref[k].abs().max()), 1e-6)
            l2 = float(d.norm()) / max(float(grads_ref[k].double().norm()), 1e-12)
            worst, worst_l2 = max(worst, e), max(worst_l2, l2)
            assert e <= 5e-3, f"{k}: {e:.3e} of max |grad|"
            assert l2 <= 3e-3, f"{k}: relative L2 error {l2:.3e}"
    print(f"bf16x3 full-size step: recon err {err:.2e}, worst grad err {worst:.2e} of max |grad|, worst rel L2 {worst_l2:.2e}")


def test_performer_config4_depth2_against_oracle_at_14000_tokens():
    """Performer of BASELINE config 4 at its full sequence length (grid 20 x 28 x 25 = 14 000 tokens, 34 local windows,
    110 scan chunks), two layers, batch 1, gates opened: fp32 CUDA-core path at 1e-4 (2e-4 of max |grad|), then the
    benchmarked bf16 tcgen05 path at a stated bound -- logits within 8 bf16 ulp of their range (2 layers x (attention +
    FFN) roundings on a residual stream kept in fp32), loss within 1e-2, gradient cosines >= 0.99."""
    from oracle import performer_oracle as po
    from synthanatomy_b200.losses import CELoss
    from tests.test_gpu_performer import _build, _close
    kw = dict(num_tokens=2049, dim=512, depth=2, heads=16, dim_head=64, local_attn_heads=8, local_window_size=420)
    grid = (20, 28, 25)
    ref = None
    for dt in (None, torch.bfloat16):
        cfg, sd, net, seqs, x_in, y = _build(kw, grid, 31, compute_dtype=dt)      # deterministic in the seed
        x_in, y = x_in[:1], y[:1]
        if ref is None:
            ref = po.train_step_grads(sd, cfg, x_in, y, seqs)
        loss_ref, grads_ref, logits_ref = ref
        net = net.cuda().train()
        logits = net(x_in.cuda())
        loss = CELoss()(logits.transpose(1, 2), y.cuda())
        loss.backward()
        named = dict(net.named_parameters())
        if dt is None:
            _close(logits, logits_ref, 1e-4, "fp32 logits @ N=14000")
            assert abs(float(loss) - float(loss_ref)) <= 1e-4 * max(1.0, abs(float(loss_ref)))
            for k, gref in grads_ref.items():
                err = float((named[k].grad.cpu() - gref).abs().max()) / max(float(gref.abs().max()), 1e-3)
                assert err <= 2e-4, f"fp32 grad {k}: {err:.3e}"
        else:
            err = float((logits.cpu() - logits_ref).abs().max()) / float(logits_ref.abs().max())
            assert err <= 8 * ULP, f"bf16 logits: {err:.3e} of max |logit|"
            assert abs(float(loss) - float(loss_ref)) <= 1e-2 * abs(float(loss_ref))
            for k, gref in grads_ref.items():
                if float(gref.abs().max()) < 1e-7:
                    continue
                a, b = named[k].grad.cpu().flatten().double(), gref.flatten().double()
                cos = float((a @ b) / (a.norm() * b.norm() + 1e-30))
                assert cos >= 0.99, f"bf16 grad {k}: cosine {cos:.4f}"
        del net
        torch.cuda.empty_cache()
