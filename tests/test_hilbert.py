"""Generalised Hilbert curve of the `hilbert_curve` ordering type: the product's iterative routine against the oracle's
recursive restatement, against the curve's defining properties, and -- where the reference tree is present -- against the
vendored generators the reference calls (img2seq_ordering.py:196-201).  No GPU."""
import itertools
import os
import sys

import numpy as np
import pytest

from oracle import performer_oracle as po
from synthanatomy_b200.networks.transformers.hilbert import hilbert_curve_indices

SHAPES = [(1, 1), (1, 7), (6, 1), (2, 2), (5, 12), (12, 5), (7, 7), (16, 16), (1, 1, 1), (2, 2, 2), (3, 7, 5), (4, 4, 9),
          (10, 14, 10), (8, 8, 8), (5, 1, 6), (1, 1, 9)]


@pytest.mark.parametrize("shape", SHAPES)
def test_curve_visits_every_cell_once_and_matches_the_oracle(shape):
    idx = hilbert_curve_indices(*shape)
    n = int(np.prod(shape))
    assert idx.shape == (n, len(shape)) and idx.min() >= 0
    assert all(idx[:, i].max() == shape[i] - 1 for i in range(len(shape)))
    flat = np.ravel_multi_index(tuple(idx.T), shape)
    assert np.array_equal(np.sort(flat), np.arange(n))                      # a permutation of the grid
    assert np.array_equal(idx, np.array(list(po.gilbert_curve(shape))).reshape(n, len(shape)))
    assert tuple(idx[0]) == (0,) * len(shape)


@pytest.mark.parametrize("shape", [(4, 4), (8, 6), (16, 16), (2, 2, 2), (4, 6, 8), (8, 8, 8), (10, 14, 10)])
def test_even_boxes_give_a_continuous_path(shape):
    """with even extents consecutive cells are face neighbours (the property that makes the ordering worth having)"""
    idx = hilbert_curve_indices(*shape)
    assert (np.abs(np.diff(idx, axis=0)).sum(axis=1) == 1).all()


def test_bad_arguments():
    for bad in [(4,), (2, 3, 4, 5), (0, 3), (3, -1, 2)]:
        with pytest.raises(ValueError):
            hilbert_curve_indices(*bad)


def test_ordering_class_accepts_hilbert_curve():
    from synthanatomy_b200.networks.transformers import Ordering
    o = Ordering("hilbert_curve", 3, (1, 4, 6, 4), (False,) * 3, ((0, 1, 2),), (), ("transpose",))
    seq = o.get_sequence_ordering()
    assert np.array_equal(np.sort(seq), np.arange(96))
    assert np.array_equal(seq[o.get_revert_sequence_ordering()], np.arange(96))
    assert np.array_equal(seq, np.ravel_multi_index(tuple(hilbert_curve_indices(4, 6, 4).T), (4, 6, 4)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/gilbert"), reason="reference tree not present (GPU box)")
def test_sweep_against_the_vendored_generators():
    sys.path.insert(0, "/root/reference")
    try:
        from gilbert.gilbert2d import gilbert2d
        from gilbert.gilbert3d import gilbert3d
    finally:
        sys.path.pop(0)
    for w, h in itertools.product(range(1, 12), repeat=2):
        assert np.array_equal(hilbert_curve_indices(w, h), np.array(list(gilbert2d(w, h)))), (w, h)
    for w, h, d in itertools.product(range(1, 7), repeat=3):
        assert np.array_equal(hilbert_curve_indices(w, h, d), np.array(list(gilbert3d(w, h, d)))), (w, h, d)
    assert np.array_equal(hilbert_curve_indices(20, 28, 25), np.array(list(gilbert3d(20, 28, 25))))
