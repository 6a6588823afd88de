"""Generate tests/golden/performer_host.npz by running the UNMODIFIED host-side reference modules of the Performer
path from /root/reference (build container only):

* ``src/networks/transformers/img2seq_ordering.py`` (imports cleanly) -> ``_sequence_ordering`` for several configs;
* ``src/utils/transformer.py:prepare_batch`` (its module imports ignite / monai at the top for the data flow, which
  are absent here: stubbed with empty modules; ``convert_tensor`` -- the only stubbed symbol prepare_batch calls --
  is ``tensor.to(device)`` and is the identity for device=None).

The Performer arithmetic itself cannot be executed here (performer-pytorch / local-attention / fast-transformers are
not installed and not vendored): see the header of oracle/performer_oracle.py ("parity unpinned").
TEST INFRASTRUCTURE -- never imported by the product.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

ORDER_CASES = [
    # name, type, dims, reflected, transpositions, rot90, order
    ("readme_10x14x10", "raster_scan", (1, 10, 14, 10), (False,) * 3, ((2, 0, 1),), ((0, 1),), ("rotate_90", "transpose")),
    ("s_curve_4x5x6", "s_curve", (1, 4, 5, 6), (False,) * 3, ((2, 0, 1),), ((0, 1),), ("rotate_90", "transpose")),
    ("reflect_3x4x5", "raster_scan", (1, 3, 4, 5), (True, False, True), ((1, 0, 2),), ((1, 2),),
     ("transpose", "rotate_90", "reflect")),
    ("s_curve_2d_4x5", "s_curve", (1, 4, 5), (False, True), ((1, 0),), ((0, 1),), ("transpose", "rotate_90", "reflect")),
    ("hilbert_10x14x10", "hilbert_curve", (1, 10, 14, 10), (False,) * 3, ((2, 0, 1),), ((0, 1),), ("rotate_90", "transpose")),
    ("hilbert_odd_3x7x5", "hilbert_curve", (1, 3, 7, 5), (False, True, False), ((0, 2, 1),), ((0, 2),),
     ("reflect", "transpose", "rotate_90")),
    ("hilbert_2d_5x12", "hilbert_curve", (1, 5, 12), (True, False), ((1, 0),), ((0, 1),), ("transpose", "rotate_90", "reflect")),
]


def main():
    for name in ("ignite", "ignite.utils", "monai", "monai.data", "monai.transforms", "monai.transforms.io",
                 "monai.transforms.io.dictionary"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["ignite.utils"].convert_tensor = lambda t, device=None, non_blocking=False: t if device is None else t.to(device)
    for n in ("Dataset", "DataLoader", "DistributedSampler"):
        setattr(sys.modules["monai.data"], n, object)
    for n in ("ToTensord", "Compose"):
        setattr(sys.modules["monai.transforms"], n, object)
    sys.modules["monai.transforms.io.dictionary"].LoadImaged = object
    sys.path.insert(0, REF)
    from src.networks.transformers.img2seq_ordering import Ordering
    from src.utils.transformer import prepare_batch

    blob = {}
    for name, typ, dims, refl, tr, rot, order in ORDER_CASES:
        o = Ordering(typ, len(dims) - 1, dims, refl, tr, rot, order)
        blob[f"order/{name}"] = o.get_sequence_ordering().astype(np.int32)
        blob[f"revert/{name}"] = o.get_revert_sequence_ordering().astype(np.int32)
    # prepare_batch on the README grid
    o = Ordering(*[ORDER_CASES[0][1], 3, *ORDER_CASES[0][2:]])
    q = torch.randint(0, 2048, (3, 10, 14, 10), generator=torch.Generator().manual_seed(2)).to(torch.int16)
    (x_in, cond), y = prepare_batch({"quantization": q}, o.get_sequence_ordering(), 2048)
    blob["pb/quantization"] = q.numpy()
    blob["pb/x_input"] = x_in.numpy().astype(np.int16)
    blob["pb/x_target"] = y.numpy().astype(np.int16)
    assert cond is None
    np.savez_compressed(os.path.join(OUT, "performer_host.npz"), **blob)
    print("performer_host written:", {k: v.shape for k, v in blob.items()})


if __name__ == "__main__":
    main()
