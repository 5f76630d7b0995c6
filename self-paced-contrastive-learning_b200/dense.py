"""Dense (pixel) contrast front end: DenseProjectionHead's tail and the dense hook's point sampling.

Reference (paths under ``/root/reference``):

* ``contrastyou/projectors/heads.py:109-115`` -- ``DenseProjectionHead.forward`` ends with
  ``AdaptiveAvgPool2d(spatial_size)`` and ``Normalize()`` (``F.normalize(dim=1)``);
* ``semi_seg/hooks/infonce.py:20-22`` -- ``get_n_point_coordinate(h, w, n)``: ``n`` distinct rows and ``n``
  distinct columns drawn with the legacy global numpy generator;
* ``semi_seg/hooks/infonce.py:233-241`` -- ``region_extractor``: per image, the feature vectors at those
  coordinates, concatenated to ``[B * point_nums, C]``;
* ``contrastyou/epocher/comparable.py:398-404`` -- the all-pixels form ``[b, c, h, w] -> [b*h*w, c]``.

Here pooling, normalisation and the gather / reshape are one CUDA pass (``ops.dense_rows``); with sampled points
only the sampled pooling windows are ever read.  ``FixRandomSeed`` comes from ``deepclustering2`` (un-pinned
git dependency, absent from the reference tree); its published behaviour is to seed ``random``, ``numpy.random``
and ``torch`` with the given seed for the duration of the block, so ``point_coordinates(seed=...)`` restates
the draw with ``numpy.random.RandomState(seed)``, whose stream is identical to ``numpy.random.seed(seed)``.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor, nn

from . import ops

__all__ = ["get_n_point_coordinate", "point_coordinates", "region_extractor", "DenseProjectionTail"]


def get_n_point_coordinate(h: int, w: int, n: int, rng=None):
    """infonce.py:20-22.  ``rng``: a ``numpy.random.RandomState`` (default: the global legacy generator,
    exactly what the reference draws from)."""
    rng = np.random if rng is None else rng
    return [(x, y) for x, y in zip(rng.choice(range(h), n, replace=False), rng.choice(range(w), n, replace=False))]


def point_coordinates(batch: int, h: int, w: int, point_nums: int = 5, seed: Optional[int] = None) -> Tensor:
    """Flat coordinates ``x * w + y`` for every image, int32 ``[batch, point_nums]`` on the host, in the order
    ``region_extractor`` draws them (image by image; rows first, then columns: infonce.py:238-241).
    ``seed`` reproduces ``with FixRandomSeed(seed): region_extractor(...)`` (infonce.py:209-212)."""
    rng = np.random if seed is None else np.random.RandomState(seed)
    out = np.empty((batch, point_nums), dtype=np.int32)
    for b in range(batch):
        for p, (x, y) in enumerate(get_n_point_coordinate(h, w, point_nums, rng)):
            out[b, p] = int(x) * w + int(y)
    return torch.from_numpy(out)


def region_extractor(features: Tensor, point_nums: int = 5, *, spatial_size=None, seed: Optional[int] = None,
                     points: Optional[Tensor] = None) -> Tensor:
    """Drop-in for ``_INFONCEDenseHook.region_extractor`` (infonce.py:233-241) on the UN-normalised, un-pooled
    projector output: returns ``[B * point_nums, C]`` unit rows equal to
    ``region_extractor(F.normalize(adaptive_avg_pool2d(features, spatial_size), dim=1), point_nums)``.
    Feeding already pooled / normalised maps (the reference's call) is the ``spatial_size=None`` case:
    normalising unit vectors again is the identity up to rounding."""
    h, w = features.shape[2:] if spatial_size is None else (
        (spatial_size, spatial_size) if isinstance(spatial_size, int) else tuple(spatial_size))
    if points is None:
        points = point_coordinates(features.shape[0], h, w, point_nums, seed)
    return ops.dense_rows(features, (h, w), points)


class DenseProjectionTail(nn.Module):
    """``pool -> normalise -> rows`` of ``DenseProjectionHead`` (heads.py:97-115) after its ``_projector`` convs.

    ``forward(out)`` returns all pooled pixels as rows ``[B*ph*pw, C]`` (what the loss consumes after
    comparable.py:398-404); ``forward(out, points=...)`` / ``forward(out, point_nums=5, seed=s)`` returns the
    sampled rows of the dense hook.  ``pool_name``: ``"adaptive_avg"`` (the reference's default, heads.py:99; the
    bandwidth-tuned kernels) or ``"adaptive_max"`` (nn.py:57-58).  ``normalize=False`` raises: the contrastive loss
    asserts unit rows (contrast_loss3.py:154)."""

    def __init__(self, spatial_size: Sequence[int] = (16, 16), pool_name: str = "adaptive_avg", normalize: bool = True):
        super().__init__()
        if pool_name not in ("adaptive_avg", "adaptive_max"):
            raise NotImplementedError(f"pool_name {pool_name!r}: the dense contrast heads pool with adaptive_avg / adaptive_max")
        if not normalize:
            raise NotImplementedError("the contrastive loss asserts unit rows (contrast_loss3.py:154)")
        self._pool = "max" if pool_name == "adaptive_max" else "avg"
        self._spatial_size: Tuple[int, int] = tuple(int(v) for v in spatial_size)

    def forward(self, out: Tensor, points: Optional[Tensor] = None, point_nums: Optional[int] = None,
                seed: Optional[int] = None) -> Tensor:
        if points is None and point_nums is not None:
            points = point_coordinates(out.shape[0], *self._spatial_size, point_nums, seed)
        return ops.dense_rows(out, self._spatial_size, points, pool=self._pool)
