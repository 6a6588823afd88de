"""GPU parity of the B200VQVAE module (plugin boundary) against the reference's own outputs in tests/golden
(BASELINE.json configs[0] and a 2-level case): forward, all parameter gradients, EMA state, eval-mode API slices."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests import golden_util as gu

pytestmark = pytest.mark.gpu


def _build(name, **kw):
    from synthanatomy_b200.networks.vqvae import B200VQVAE
    cfg, sd, blob = gu.vqvae_case(name)
    net = B200VQVAE(**cfg, **kw)
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return net.cuda(), blob


@pytest.mark.parametrize("name", ["vqvae_cfg1", "vqvae_l2"])
def test_fp32_training_step_matches_reference(name):
    net, blob = _build(name)
    net.train()
    x = torch.from_numpy(blob["x"]).cuda()
    out = net(x)
    recon, q_loss = out["reconstruction"][0], out["quantization_losses"][0]
    loss = F.mse_loss(recon.float(), x.float()) + q_loss.float()
    loss.backward()
    # north_star tolerance: 1e-4 in fp32
    np.testing.assert_allclose(recon.detach().cpu().numpy(), blob["recon"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(q_loss.detach().cpu().numpy(), blob["q_loss"], rtol=1e-4)
    np.testing.assert_allclose(loss.detach().cpu().numpy(), blob["loss"], rtol=1e-4)
    np.testing.assert_allclose(net.get_perplexity()[0].cpu().numpy(), blob["perplexity"], rtol=1e-4)
    for k, p in net.named_parameters():
        if not p.requires_grad:
            assert p.grad is None
            continue
        g = blob["grad/" + k]
        scale = max(np.abs(g).max(), 1e-6)
        np.testing.assert_allclose(p.grad.cpu().numpy(), g, rtol=1e-3, atol=1e-4 * scale, err_msg=k)
    sd1 = net.state_dict()
    for k in ("quantizer.0.impl.weight", "quantizer.0.impl.N", "quantizer.0.impl.embed_avg"):
        np.testing.assert_allclose(sd1[k].cpu().numpy(), blob["sd1/" + k], rtol=1e-4, atol=1e-6, err_msg=k)
    # aliasing of the codebook parameter (baseline.py:33)
    assert sd1["quantizer.0.impl.weight"].data_ptr() == sd1["quantizer.0.impl.embedding.weight"].data_ptr()


@pytest.mark.parametrize("name", ["vqvae_cfg1", "vqvae_l2"])
def test_eval_api_slices(name):
    net, blob = _build(name)
    with torch.no_grad():
        net.quantizer[0].impl.embedding.weight.copy_(torch.from_numpy(blob["sd1/quantizer.0.impl.weight"]))
    net.eval()
    x = torch.from_numpy(blob["x"]).cuda()
    with torch.no_grad():
        enc = net.encode(x)[0]
        idx = net.index_quantize(x)[0]
        dec = net.decode_samples([idx])
    np.testing.assert_allclose(enc.cpu().numpy(), blob["eval_encode"], rtol=1e-4, atol=1e-4)
    assert idx.dtype == torch.int64
    np.testing.assert_array_equal(idx.cpu().numpy(), blob["eval_idx"])        # bit-exact indices
    np.testing.assert_allclose(dec.cpu().numpy(), blob["eval_decode"], rtol=1e-4, atol=1e-4)
    # decoding the reference's own indices (decoding mode, src/inferer/vqvae.py:82)
    dec2 = net.decode_samples([torch.from_numpy(blob["eval_idx"]).cuda()])
    np.testing.assert_allclose(dec2.detach().cpu().numpy(), blob["eval_decode"], rtol=1e-4, atol=1e-4)


def test_getters_setters_and_last_layer():
    net, _ = _build("vqvae_l2")
    assert net.get_ema_decay() == [0.5] and net.set_ema_decay([0.9]) == [0.9] and net.set_ema_decay(0.7) == [0.7]
    assert net.get_commitment_cost() == [0.25] and net.set_commitment_cost(0.5) == [0.5]
    last = net.get_last_layer()
    assert last is net.decoder[0][-1].weight and last.shape[1] == 1


def test_bf16_tensor_core_step_tracks_fp32_oracle():
    """Throughput mode (bf16 operands, fp32 accumulate; the reference's --amp analogue) on a shape wide enough for
    the tcgen05 kernels, against the CPU oracle in fp32.  Tolerance: bf16 has 8 mantissa bits; through ~20 stacked
    convs the reconstruction error stays within a few 1e-2 of the output range."""
    from oracle import vqvae_oracle as vo
    from synthanatomy_b200 import ops
    from synthanatomy_b200.networks.vqvae import B200VQVAE
    kw = dict(n_levels=2, downsample_parameters=((4, 2, 1, 1),) * 2, upsample_parameters=((4, 2, 1, 0, 1),) * 2,
              n_embed=64, embed_dim=32, n_channels=256, n_res_channels=256, n_res_layers=1, vq_decay=0.5,
              commitment_cost=0.25)
    torch.manual_seed(1)
    net = B200VQVAE(**kw)
    with torch.no_grad():
        net.quantizer[0].impl.embedding.weight.mul_(0.05)
        net.quantizer[0].impl.embed_avg.copy_(net.quantizer[0].impl.embedding.weight)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    x = torch.rand(1, 1, 32, 32, 32)
    cfg = vo.VQVAEConfig(**kw)
    with torch.no_grad():
        z_ref = vo.encode(sd, cfg, x)
    net = net.cuda().train()
    ops.reset_launch_count()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        z = net.encode(x.cuda())[0]
    assert ops.last_path() == 2, "tcgen05 path not taken"
    err = (z.cpu() - z_ref).abs().max() / z_ref.abs().max()
    assert err < 5e-2, f"bf16 encoder deviates {err:.3e} from the fp32 oracle"
    # full step runs and produces finite grads of the right shapes
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = net(x.cuda())
    loss = F.mse_loss(out["reconstruction"][0].float(), x.cuda()) + out["quantization_losses"][0]
    loss.backward()
    for k, p in net.named_parameters():
        if p.requires_grad:
            assert p.grad is not None and torch.isfinite(p.grad).all(), k
    assert ops.launch_count() > 50


def test_bf16_grads_close_to_oracle():
    """1-level, 128 channels: exercises the single-channel im2col / col2im GEMM paths (first conv 1->128, last
    transposed conv 128->1) and the tcgen05 conv / dgrad / wgrad kernels end to end; every parameter gradient must
    point the same way as the fp32 oracle's (cosine > 0.99) and agree to a few bf16 ulps of its scale."""
    from oracle import vqvae_oracle as vo
    from synthanatomy_b200.networks.vqvae import B200VQVAE
    kw = dict(n_levels=1, downsample_parameters=((4, 2, 1, 1),), upsample_parameters=((4, 2, 1, 0, 1),),
              n_embed=64, embed_dim=16, n_channels=128, n_res_channels=128, n_res_layers=2, vq_decay=0.5,
              commitment_cost=0.25)
    torch.manual_seed(3)
    net = B200VQVAE(**kw)
    with torch.no_grad():
        net.quantizer[0].impl.embedding.weight.mul_(0.05)
        net.quantizer[0].impl.embed_avg.copy_(net.quantizer[0].impl.embedding.weight)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    x = torch.rand(2, 1, 16, 24, 32)
    loss_ref, grads_ref, out_ref = vo.train_step_grads(sd, vo.VQVAEConfig(**kw), x)
    net = net.cuda().train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = net(x.cuda())
    loss = F.mse_loss(out["reconstruction"][0].float(), x.cuda()) + out["quantization_losses"][0]
    loss.backward()
    rerr = (out["reconstruction"][0].cpu() - out_ref["reconstruction"][0]).abs().max().item()
    assert rerr < 5e-2, rerr
    assert abs(loss.item() - loss_ref.item()) < 2e-2 * abs(loss_ref.item())
    for k, p in net.named_parameters():
        if not p.requires_grad:
            continue
        a, b = p.grad.cpu().flatten().double(), grads_ref[k].flatten().double()
        cos = float((a @ b) / (a.norm() * b.norm() + 1e-30))
        assert cos > 0.99, f"{k}: cosine {cos:.4f}"
        assert float((a - b).abs().max()) < 0.1 * float(b.abs().max()) + 1e-7, k


def test_amp_gradscaler_iteration_as_the_reference_trainer_runs_it():
    """the generator half of AdversarialTrainer._iteration (/root/reference/src/engines/trainer.py:157-190) /
    MONAI's SupervisedTrainer with amp=True: forward under torch.cuda.amp.autocast() (fp16 autocast in the reference; the
    drop-in answers any autocast with its bf16 tensor-core path), GradScaler.scale(loss).backward(), scaler.step,
    scaler.update -- with stock torch.optim.Adam as run_vqvae.py:82 builds it."""
    from synthanatomy_b200.networks.vqvae import B200VQVAE
    from synthanatomy_b200 import ops
    kw = dict(n_levels=1, downsample_parameters=((4, 2, 1, 1),), upsample_parameters=((4, 2, 1, 0, 1),),
              n_embed=64, embed_dim=16, n_channels=128, n_res_channels=128, n_res_layers=2, vq_decay=0.5,
              commitment_cost=0.25)
    torch.manual_seed(0)
    net = B200VQVAE(**kw).cuda().train()
    opt = torch.optim.Adam(net.parameters(), 1.65e-4)
    scaler = torch.amp.GradScaler("cuda")
    x = torch.rand(2, 1, 16, 16, 32, device="cuda")
    before = [p.detach().clone() for p in net.parameters()]
    losses = []
    for _ in range(3):
        opt.zero_grad()
        with torch.autocast("cuda"):                      # default dtype of CUDA autocast = float16, as in the reference
            out = net(x)
            loss = F.mse_loss(out["reconstruction"][0].float(), x) + out["quantization_losses"][0]
        assert ops.last_path() == 2, "the tensor-core path did not answer the autocast region"
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        losses.append(float(loss))
    assert all(torch.isfinite(p).all() for p in net.parameters())
    assert sum(int(not torch.equal(a, p)) for a, p in zip(before, net.parameters())) >= len(before) - 2   # codebook: EMA
    assert scaler.get_scale() > 0 and losses[-1] < losses[0]
