"""CPU, world size 2 over gloo: the host-side logic of the N > 1 path.

The product computes on the GPU only, so what can be exercised here is the distributed protocol itself:
  * the packed (counts | dw) all-reduce of the quantiser statistics (b200.py:_QuantizeFn, reference baseline.py:70-72):
    per-rank statistics of a batch shard, summed over ranks, give the single-process EMA update -- checked with the
    oracle's arithmetic on both sides;
  * bench.py --impl reference under a 2-rank launch prints exactly one JSON line (rank 0) and the other rank exits 0.
"""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from oracle import vqvae_oracle as vo
    K, D = 64, 8
    g = torch.Generator().manual_seed(0)
    W = torch.randn(K, D, generator=g) * 0.5
    z = torch.randn(4, D, 2, 3, 2, generator=g)                  # global batch of 4, 2 per rank
    cfg = vo.VQVAEConfig(n_levels=1, downsample_parameters=((4, 2, 1, 1),), upsample_parameters=((4, 2, 1, 0, 1),),
                         n_embed=K, embed_dim=D, n_channels=8, n_res_channels=8)
    sd = {"quantizer.0.impl.weight": W, "quantizer.0.impl.N": torch.zeros(K), "quantizer.0.impl.embed_avg": W.clone()}

    def packed_all_reduce():
        """what b200.py:_QuantizeFn does: ONE collective over the packed (counts | dw) buffer"""
        box = {}

        def hook(t):
            if t.dim() == 1:                       # encodings_sum arrives first (baseline.py:71), dw second (:72)
                box["counts"] = t
                return t
            stats = torch.cat([box["counts"], t.reshape(-1)])
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
            box["counts"].copy_(stats[:K])
            return stats[K:].view(K, D)
        return hook

    shard = z[rank * 2:(rank + 1) * 2]
    _, _, _, st = vo.quantize(sd, cfg, shard, True, all_reduce=packed_all_reduce())
    _, _, _, st1 = vo.quantize(sd, cfg, z, True)                 # single process, whole batch
    new_w = st["weight"]
    ok = all(torch.allclose(st[k], st1[k], atol=1e-6) for k in ("N", "embed_avg", "weight"))
    # every rank ends with the same codebook
    gathered = [torch.zeros_like(new_w) for _ in range(world)]
    dist.all_gather(gathered, new_w)
    ok = ok and all(torch.equal(gathered[0], t) for t in gathered)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_packed_ema_statistics_allreduce_world2():
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29611, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0] and ret[1]


def test_reference_arm_prints_once_under_two_ranks():
    outs = []
    for rank in (0, 1):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT="29612")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "performer",
                            "--pf-depth", "1", "--steps", "1", "--warmup", "0", "--gpus", "2"], env=env,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout.strip())
    line = json.loads(outs[0])
    assert line["impl"] == "reference" and line["unit"] == "tokens/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert outs[1] == ""
