"""Generate tests/golden/jukebox.npz by running the UNMODIFIED reference ``JukeboxLoss``
(/root/reference/src/losses/vqvae/vqvae.py:522-640) on CPU.  Build container only.

    python oracle/make_golden_losses.py

Stubs: ``lpips.LPIPS`` (imported at vqvae.py:6, used only by the perceptual losses) and ``src.handlers.general``
(imports Ignite / MONAI at module level; only its ``TBSummaryTypes`` enum is touched here).
TEST INFRASTRUCTURE -- never imported by the product.
"""
from __future__ import annotations

import enum
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "jukebox.npz")


def main():
    lp = types.ModuleType("lpips")
    lp.LPIPS = type("LPIPS", (), {})
    sys.modules["lpips"] = lp
    hg = types.ModuleType("src.handlers.general")

    class TBSummaryTypes(enum.Enum):
        SCALAR = "scalar"

    hg.TBSummaryTypes = TBSummaryTypes
    sys.path.insert(0, REF)
    import src                                   # noqa: F401  (namespace of the reference tree)
    sys.modules["src.handlers"] = types.ModuleType("src.handlers")
    sys.modules["src.handlers.general"] = hg
    from src.losses.vqvae.vqvae import JukeboxLoss

    g = torch.Generator().manual_seed(21)
    y = torch.rand(2, 1, 8, 12, 10, generator=g)
    pred = (y + 0.1 * torch.randn(2, 1, 8, 12, 10, generator=g)).requires_grad_(True)
    q = torch.tensor(0.0375)
    store = {"y": y.numpy().copy(), "pred": pred.detach().numpy().copy(), "q": q.numpy().copy()}
    for name, kw, factor in (("default", {}, 1.0), ("no_pixel_f2", {"include_pixel_loss": False}, 2.0)):
        loss_fn = JukeboxLoss(dimensions=3, **kw)
        loss_fn.set_fft_factor(factor)
        pred.grad = None
        loss = loss_fn({"reconstruction": [pred], "quantization_losses": [q]}, y)
        loss.backward()
        store[f"{name}/loss"] = np.float64(loss.item())
        store[f"{name}/grad"] = pred.grad.numpy().copy()
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes", {k: float(v) for k, v in store.items() if k.endswith("loss")})


if __name__ == "__main__":
    main()
