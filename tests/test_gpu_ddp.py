"""DistributedDataParallel over NCCL exactly as the reference's entry points wrap the two networks
(/root/reference/run_vqvae.py:71-77: broadcast_buffers=False, bucket_cap_mb=12.5;
 /root/reference/run_transformer.py:98-105: broadcast_buffers=True, find_unused_parameters=True, bucket_cap_mb=12.5).

  * one GPU, world size 1: the wrapped drop-ins train and produce the gradients of the bare modules (the reducer's
    hooks, unused-parameter search and buffer broadcast all run against the hand-scheduled autograd Functions);
  * two GPUs (skipped on a 1-GPU box; run with `gpurun --gpus 2`): the 2-rank run over a split batch reproduces the
    1-GPU run over the whole batch -- same loss, same gradients, and the EMA codebook is identical on both ranks and
    equal to the single-GPU one ("distributed").
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

VQ_KW = dict(n_levels=2, downsample_parameters=((4, 2, 1, 1),) * 2, upsample_parameters=((4, 2, 1, 0, 1),) * 2,
             n_embed=64, embed_dim=16, n_channels=128, n_res_channels=128, n_res_layers=1, vq_decay=0.5,
             commitment_cost=0.25)
PF_KW = dict(num_tokens=65, dim=128, depth=2, heads=4, dim_head=64, local_attn_heads=2, local_window_size=20)
PF_GRID = (4, 5, 6)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _vqvae(seed=0):
    from synthanatomy_b200.networks.vqvae import B200VQVAE
    torch.manual_seed(seed)
    net = B200VQVAE(**VQ_KW)
    with torch.no_grad():
        net.quantizer[0].impl.embedding.weight.mul_(0.05)
        net.quantizer[0].impl.embed_avg.copy_(net.quantizer[0].impl.embedding.weight)
    return net


def _performer(seed=0):
    from synthanatomy_b200.networks.transformers import Ordering, Performer
    torch.manual_seed(seed)
    n = int(np.prod(PF_GRID))
    order = Ordering("raster_scan", 3, (1, *PF_GRID), (False,) * 3, ((2, 0, 1),), ((0, 1),), ("rotate_90", "transpose"))
    net = Performer(max_seq_len=n + 1, ordering=order, causal=True, feature_redraw_interval=1, use_rezero=True,
                    spatial_position_emb="absolute", spatial_shape=PF_GRID, conditioning_num_tokens=0, **PF_KW)
    with torch.no_grad():
        for layer in net.performer.net.layers:
            layer[0].g.fill_(0.7); layer[1].g.fill_(-0.4)
    return net, n


def _vq_step(model, x):
    out = model(x)
    loss = F.mse_loss(out["reconstruction"][0], x) + out["quantization_losses"][0]
    loss.backward()
    return loss.detach()


def _pf_step(model, tok, tgt):
    logits = model(tok)                                   # TransformerTrainingInferer: network(seq).transpose(1, 2)
    loss = F.cross_entropy(logits.transpose(1, 2), tgt)
    loss.backward()
    return loss.detach()


def test_reference_ddp_wrapping_single_rank():
    port = _free_port()
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1,
                            device_id=torch.device("cuda", 0))
    try:
        # ---- VQ-VAE, run_vqvae.py:71-77
        bare, net = _vqvae().cuda().train(), _vqvae().cuda().train()
        ddp = torch.nn.parallel.DistributedDataParallel(net, device_ids=[0], broadcast_buffers=False, bucket_cap_mb=12.5)
        x = torch.rand(2, 1, 16, 16, 16, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
        for it in range(2):                                # twice: the reducer rebuilds its buckets after the first step
            bare.zero_grad(); ddp.zero_grad()
            l0, l1 = _vq_step(bare, x), _vq_step(ddp, x)
            # (not bit-equal: the statistics / split-K reductions use fp32 atomics, whose order varies run to run)
            assert abs(float(l0) - float(l1)) <= 1e-6 * abs(float(l0)), (it, float(l0), float(l1))
            for (k, p), (_, q) in zip(bare.named_parameters(), net.named_parameters()):
                if p.requires_grad:
                    assert q.grad is not None, k
                    torch.testing.assert_close(q.grad, p.grad, rtol=1e-4, atol=1e-6 * float(p.grad.abs().max() + 1), msg=k)
        torch.testing.assert_close(bare.quantizer[0].impl.weight, net.quantizer[0].impl.weight, rtol=1e-5, atol=1e-6)
        # ---- Performer, run_transformer.py:98-105 (find_unused_parameters=True, broadcast_buffers=True)
        (bare, n), (net, _) = _performer(), _performer()
        bare, net = bare.cuda().train(), net.cuda().train()
        ddp = torch.nn.parallel.DistributedDataParallel(net, device_ids=[0], broadcast_buffers=True,
                                                        find_unused_parameters=True, bucket_cap_mb=12.5)
        g = torch.Generator(device="cuda").manual_seed(2)
        tok = torch.randint(0, 64, (2, n), device="cuda", generator=g)
        tgt = torch.randint(0, 64, (2, n), device="cuda", generator=g)
        for it in range(2):
            net.load_state_dict(bare.state_dict())         # same projection buffers on both sides
            bare.performer.proj_updater._calls = net.performer.proj_updater._calls = 0
            bare.zero_grad(); ddp.zero_grad()
            l0, l1 = _pf_step(bare, tok, tgt), _pf_step(ddp, tok, tgt)
            assert abs(float(l0) - float(l1)) <= 1e-6 * abs(float(l0)), (it, float(l0), float(l1))
            for (k, p), (_, q) in zip(bare.named_parameters(), net.named_parameters()):
                assert q.grad is not None, k
                torch.testing.assert_close(q.grad, p.grad, rtol=1e-5, atol=1e-7, msg=k)
    finally:
        dist.destroy_process_group()


def _two_rank_worker(rank, port, out_dir):
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=2,
                            device_id=torch.device("cuda", rank))
    try:
        dev = torch.device("cuda", rank)
        x = torch.rand(4, 1, 16, 16, 16, generator=torch.Generator().manual_seed(1))          # the GLOBAL batch
        net = _vqvae().to(dev).train()
        ddp = torch.nn.parallel.DistributedDataParallel(net, device_ids=[rank], broadcast_buffers=False, bucket_cap_mb=12.5)
        loss = _vq_step(ddp, x[2 * rank: 2 * rank + 2].to(dev))
        res = {"vq_loss": loss.cpu(), "vq_codebook": net.quantizer[0].impl.weight.detach().cpu(),
               "vq_N": net.quantizer[0].impl.N.cpu(),
               "vq_grads": {k: p.grad.cpu() for k, p in net.named_parameters() if p.requires_grad}}
        pf, n = _performer()
        pf = pf.to(dev).train()
        pf.fix_projection_matrices_()
        ddp = torch.nn.parallel.DistributedDataParallel(pf, device_ids=[rank], broadcast_buffers=True,
                                                        find_unused_parameters=True, bucket_cap_mb=12.5)
        g = torch.Generator().manual_seed(2)
        tok, tgt = torch.randint(0, 64, (4, n), generator=g), torch.randint(0, 64, (4, n), generator=g)
        loss = _pf_step(ddp, tok[2 * rank: 2 * rank + 2].to(dev), tgt[2 * rank: 2 * rank + 2].to(dev))
        res.update({"pf_loss": loss.cpu(), "pf_grads": {k: p.grad.cpu() for k, p in pf.named_parameters()}})
        torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpu_ddp_equals_single_gpu_run_over_the_global_batch(tmp_path):
    mp.spawn(_two_rank_worker, args=(_free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(tmp_path / f"rank{r}.pt", weights_only=False) for r in (0, 1))
    # ---- single GPU over the whole batch
    x = torch.rand(4, 1, 16, 16, 16, generator=torch.Generator().manual_seed(1)).cuda()
    net = _vqvae().cuda().train()
    loss = _vq_step(net, x)
    # the mean of the two per-rank losses is the global-batch loss (equal shard sizes); gradients are averaged by DDP
    assert abs(0.5 * float(r0["vq_loss"] + r1["vq_loss"]) - float(loss)) <= 1e-5 * abs(float(loss))
    assert torch.equal(r0["vq_codebook"], r1["vq_codebook"]) and torch.equal(r0["vq_N"], r1["vq_N"])
    torch.testing.assert_close(r0["vq_codebook"], net.quantizer[0].impl.weight.detach().cpu(), rtol=1e-5, atol=1e-6)
    for k, p in net.named_parameters():
        if p.requires_grad:
            assert torch.equal(r0["vq_grads"][k], r1["vq_grads"][k]), k
            # the commitment-loss gradient is local to a shard but normalised by the shard's element count: the DDP
            # average over equal shards is the global-batch value
            torch.testing.assert_close(r0["vq_grads"][k], p.grad.cpu(), rtol=1e-4, atol=1e-6 * float(p.grad.abs().max() + 1), msg=k)
    # ---- Performer: the FAVOR+ key stabiliser is a max over the PER-RANK batch in the reference (no cross-rank reduce),
    # so the 2-rank run equals the mean of two independent half-batch runs, not the full-batch run; checked as such
    pf, n = _performer()
    pf = pf.cuda().train()
    pf.fix_projection_matrices_()
    g = torch.Generator().manual_seed(2)
    tok, tgt = torch.randint(0, 64, (4, n), generator=g).cuda(), torch.randint(0, 64, (4, n), generator=g).cuda()
    losses, grads = [], None
    for r in (0, 1):
        pf.zero_grad()
        losses.append(_pf_step(pf, tok[2 * r: 2 * r + 2], tgt[2 * r: 2 * r + 2]))
        cur = {k: p.grad.clone() for k, p in pf.named_parameters()}
        grads = cur if grads is None else {k: 0.5 * (grads[k] + cur[k]) for k in cur}
    assert abs(float(r0["pf_loss"]) - float(losses[0])) <= 1e-6 * abs(float(losses[0]))
    assert abs(float(r1["pf_loss"]) - float(losses[1])) <= 1e-6 * abs(float(losses[1]))
    for k, gk in grads.items():
        assert torch.equal(r0["pf_grads"][k], r1["pf_grads"][k]), k
        torch.testing.assert_close(r0["pf_grads"][k], gk.cpu(), rtol=1e-4, atol=1e-6 * float(gk.abs().max() + 1), msg=k)
