"""Generalised Hilbert ("gilbert") space-filling curve over arbitrary rectangles / cuboids, the order behind the reference's
``hilbert_curve`` ordering type (/root/reference/src/networks/transformers/img2seq_ordering.py:196-201, which calls the
third-party ``gilbert2d`` / ``gilbert3d`` generators of J. Cerveny, BSD-2-Clause, vendored next to the reference).

Restated here as ONE iterative routine for both ranks: a box is (corner, axis vectors); it is either a one-voxel-thick line
(emitted directly) or split along the published case analysis into 2 / 3 / 5 sub-boxes that are pushed on an explicit stack.
Produces exactly the reference's visiting order (``tests/test_hilbert.py`` compares against golden sequences generated
from the vendored generators and, where the reference tree is present, against the generators themselves)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def _len(v: np.ndarray) -> int:
    return abs(int(v.sum()))            # axis vectors have a single non-zero component


def _half(v: np.ndarray, unit: np.ndarray, full: int) -> np.ndarray:
    """floor-halved axis vector, nudged to an even length when the box is longer than 2 ("prefer even steps")"""
    h = v // 2                           # floor division, also for negative components
    if _len(h) % 2 and full > 2:
        h = h + unit
    return h


def _split2(p, a, b):
    w, h = _len(a), _len(b)
    da, db = np.sign(a), np.sign(b)
    if 2 * w > 3 * h:                    # long box: two parts along the major axis
        a2 = _half(a, da, w)
        return [(p, a2, b), (p + a2, a - a2, b)]
    a2 = a // 2
    b2 = _half(b, db, h)                 # up, across, down
    return [(p, b2, a2), (p + b2, a, b - b2), (p + (a - da) + (b2 - db), -b2, -(a - a2))]


def _split3(p, a, b, c):
    w, h, d = _len(a), _len(b), _len(c)
    da, db, dc = np.sign(a), np.sign(b), np.sign(c)
    a2, b2, c2 = _half(a, da, w), _half(b, db, h), _half(c, dc, d)
    if 2 * w > 3 * h and 2 * w > 3 * d:  # wide: split the major axis only
        return [(p, a2, b, c), (p + a2, a - a2, b, c)]
    if 3 * h > 4 * d:                    # flat: leave the third axis whole
        return [(p, b2, c, a2), (p + b2, a, b - b2, c), (p + (a - da) + (b2 - db), -b2, c, -(a - a2))]
    if 3 * d > 4 * h:                    # tall: leave the second axis whole
        return [(p, c2, a2, b), (p + c2, a, b, c - c2), (p + (a - da) + (c2 - dc), -c2, -(a - a2), b)]
    return [(p, b2, c2, a2),             # regular: five octant-like parts
            (p + b2, c, a2, b - b2),
            (p + (b2 - db) + (c - dc), a, -b2, -(c - c2)),
            (p + (a - da) + b2 + (c - dc), -c, -(a - a2), b - b2),
            (p + (a - da) + (b2 - db), -b2, c2, -(a - a2))]


def hilbert_curve_indices(*shape: int) -> np.ndarray:
    """visiting order of the curve over a grid of ``shape`` (2 or 3 extents): int array [prod(shape), len(shape)]"""
    rank = len(shape)
    if rank not in (2, 3) or min(shape) < 1:
        raise ValueError(f"hilbert_curve_indices takes 2 or 3 positive extents, got {shape}")
    eye = np.eye(rank, dtype=np.int64)
    # the longest extent leads (ties resolved towards the first axis); the remaining axes keep their order
    lead = max(range(rank), key=lambda i: (shape[i], -i))
    axes = [lead] + [i for i in range(rank) if i != lead]
    box = (np.zeros(rank, dtype=np.int64), *[eye[i] * shape[i] for i in axes])
    out: List[np.ndarray] = []
    stack = [box]
    while stack:
        p, *vecs = stack.pop()
        lens = [_len(v) for v in vecs]
        thick = [i for i, n in enumerate(lens) if n != 1]
        if len(thick) <= 1:              # a line (or a single voxel): walk it
            i = thick[0] if thick else 0
            step = np.sign(vecs[i])
            out.append(p + np.arange(lens[i], dtype=np.int64)[:, None] * step)
            continue
        parts = _split2(p, *vecs) if rank == 2 else _split3(p, *vecs)
        stack.extend(reversed(parts))
    return np.concatenate(out, axis=0)
