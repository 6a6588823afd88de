#!/usr/bin/env python
"""One training step of a bench workload between cudaProfilerStart / Stop, for launch lists:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \\
        --log-file gpurun_out/launches.csv python tools/step_profile.py {vqvae|performer} {bf16|bf16x3|fp32} [--batch B]
    python tools/launch_summary.py gpurun_out/launches.csv
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", choices=["vqvae", "performer"])
    ap.add_argument("dtype", choices=["bf16", "bf16x3", "fp32"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--depth", type=int, default=24)
    ap.add_argument("--grid", type=int, nargs=3, default=[20, 28, 25])
    ap.add_argument("--warm", type=int, default=2)
    ap.add_argument("--deterministic", action="store_true")
    a = ap.parse_args()
    from synthanatomy_b200 import ops
    if a.deterministic:
        ops.set_deterministic(True)
    dt = {"bf16": torch.bfloat16, "bf16x3": ops.BF16X3, "fp32": torch.float32}[a.dtype]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    if a.workload == "vqvae":
        from synthanatomy_b200.losses import MSELoss
        from synthanatomy_b200.networks.vqvae import B200VQVAE
        from synthanatomy_b200.optim import Adam
        torch.manual_seed(4)
        net = B200VQVAE(**bench.KW, compute_dtype=dt).to(dev).train()
        opt = Adam(net.parameters(), lr=1.65e-4)
        crit = MSELoss()
        x = torch.rand(a.batch or 8, 1, 160, 224, 160, device=dev)

        def step():
            loss = crit(net(x), x)
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
    else:
        args = argparse.Namespace(pf_batch=a.batch or 6, pf_depth=a.depth)
        S = bench._pf_setup(args, tuple(a.grid), 1, 0, 0, dev, dt)

        def step():
            S["step"](S["x_dev"], S["y_dev"])
    for _ in range(a.warm):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
