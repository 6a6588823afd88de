"""Generate tests/golden/discriminator.npz by running the UNMODIFIED reference ``BaselineDiscriminator``
(/root/reference/src/networks/discriminator/baseline.py) on CPU.  Build container only.

    python oracle/make_golden_discriminator.py

TEST INFRASTRUCTURE -- never imported by the product.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "discriminator.npz")


def main():
    sys.path.insert(0, REF)
    from src.networks.discriminator.baseline import BaselineDiscriminator
    torch.manual_seed(11)
    net = BaselineDiscriminator(input_nc=1, ndf=8, n_layers=3).train()
    store = {f"init/{k}": v.detach().numpy().copy() for k, v in net.state_dict().items()}
    x = torch.rand(2, 1, 32, 32, 32, requires_grad=True)
    out = net(x)
    # least-squares GAN style objective on the patch map (trainer.py:215-256 applies a criterion to this map)
    loss = ((out - 1.0) ** 2).mean()
    loss.backward()
    store["x"] = x.detach().numpy().copy()
    store["out"] = out.detach().numpy().copy()
    store["loss"] = np.float64(loss.item())
    store["dx"] = x.grad.numpy().copy()
    for k, p in net.named_parameters():
        store[f"grad/{k}"] = p.grad.numpy().copy()
    for k, v in net.state_dict().items():
        if "running" in k or "num_batches" in k:
            store[f"after/{k}"] = v.detach().numpy().copy()
    net.eval()
    with torch.no_grad():
        store["out_eval"] = net(x).numpy().copy()
    # the same step in float64 (the reference class, .double()): what an fp32 implementation is measured against, so
    # that the rounding of torch's own fp32 CPU kernels does not count against it
    torch.manual_seed(11)
    net64 = BaselineDiscriminator(input_nc=1, ndf=8, n_layers=3).double().train()
    x64 = x.detach().double().requires_grad_(True)
    out64 = net64(x64)
    ((out64 - 1.0) ** 2).mean().backward()
    store["f64/out"] = out64.detach().numpy().copy()
    store["f64/dx"] = x64.grad.numpy().copy()
    for k, p in net64.named_parameters():
        store[f"f64/grad/{k}"] = p.grad.numpy().copy()
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", "out shape", tuple(out.shape))


if __name__ == "__main__":
    main()
