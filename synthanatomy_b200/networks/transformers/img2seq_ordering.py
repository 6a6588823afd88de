"""Host-side mirror of /root/reference/src/networks/transformers/img2seq_ordering.py:24-201 (same constructor, same
methods, same resulting index sequences -- pinned against the reference in tests/golden/performer_host.npz).  The
reference's callers build an ``Ordering`` and hand it to ``Performer(ordering=...)`` (run_transformer.py:58-66); any
object with this interface works, including the reference's own class."""
from __future__ import annotations

from enum import Enum
from typing import Sequence, Tuple

import numpy as np
import torch


class OrderingType(Enum):
    RASTER_SCAN = "raster_scan"
    S_CURVE = "s_curve"
    RANDOM = "random"
    HILBERT = "hilbert_curve"


class OrderingTransformations(Enum):
    ROTATE_90 = "rotate_90"
    TRANSPOSE = "transpose"
    REFLECT = "reflect"


class Ordering:
    def __init__(self, ordering_type: str, spatial_dims: int, dimensions: Sequence[int],
                 reflected_spatial_dims: Sequence[bool], transpositions_axes: Sequence[Sequence[int]],
                 rot90_axes: Sequence[Sequence[int]],
                 transformation_order: Tuple[str, ...] = (OrderingTransformations.TRANSPOSE.value,
                                                          OrderingTransformations.ROTATE_90.value,
                                                          OrderingTransformations.REFLECT.value)):
        kinds = [e.value for e in OrderingType]
        assert ordering_type in kinds, f"ordering_type must be one of the following {kinds}, but got {ordering_type}."
        assert len(dimensions) == spatial_dims + 1, f"Dimensions must have length {spatial_dims + 1}."
        if len(set(transformation_order)) != len(transformation_order):
            raise ValueError(f"No duplicates are allowed. Received {transformation_order}.")
        valid = [t.value for t in OrderingTransformations]
        for tr in transformation_order:
            if tr not in valid:
                raise ValueError(f"Valid transformations are {valid} but received {tr}.")
        self.ordering_type = ordering_type
        self.spatial_dims = spatial_dims
        self.dimensions = dimensions
        self.reflected_spatial_dims = reflected_spatial_dims
        self.transpositions_axes = transpositions_axes
        self.rot90_axes = rot90_axes
        self.transformation_order = transformation_order
        spatial = tuple(dimensions[1:])
        template = np.arange(int(np.prod(spatial))).reshape(*spatial)
        for tr in transformation_order:
            if tr == OrderingTransformations.TRANSPOSE.value:
                for axes in transpositions_axes:
                    template = np.transpose(template, axes=axes)
            elif tr == OrderingTransformations.ROTATE_90.value:
                for axes in rot90_axes:
                    template = np.rot90(template, axes=axes)
            else:
                for axis, flag in enumerate(reflected_spatial_dims):
                    template = np.flip(template, axis=axis) if flag else template
        self.template = template
        self._sequence_ordering = self._read_out(np.ascontiguousarray(template))
        self._revert_sequence_ordering = np.argsort(self._sequence_ordering)

    def _read_out(self, template: np.ndarray) -> np.ndarray:
        if self.ordering_type == OrderingType.RASTER_SCAN.value:
            return template.reshape(-1).copy()
        if self.ordering_type == OrderingType.S_CURVE.value:
            t = template.copy()
            if t.ndim == 3:
                t[:, 1::2, :] = t[:, 1::2, ::-1].copy()     # depth runs backwards on odd columns
            t[1::2] = t[1::2, ::-1].copy()                  # columns run backwards on odd rows
            return t.reshape(-1)
        if self.ordering_type == OrderingType.RANDOM.value:
            idx = np.indices(template.shape).reshape(template.ndim, -1).T.copy()
            np.random.shuffle(idx)
            return np.array([template[tuple(e)] for e in idx])
        from .hilbert import hilbert_curve_indices          # generalised Hilbert curve (img2seq_ordering.py:196-201)
        idx = hilbert_curve_indices(*template.shape)
        return template[tuple(idx.T)].copy()

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        return x[self._sequence_ordering]

    def get_sequence_ordering(self) -> np.ndarray:
        return self._sequence_ordering

    def get_revert_sequence_ordering(self) -> np.ndarray:
        return self._revert_sequence_ordering
