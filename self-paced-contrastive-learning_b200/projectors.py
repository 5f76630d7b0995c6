"""Projector tail: L2 normalisation as one CUDA kernel.

Drop-in for ``contrastyou/projectors/nn.py::Normalize`` (:29-36, ``F.normalize(input, p=2, dim)``)
as used at the end of ``ProjectionHead`` (heads.py:17, ``[B, C]``, dim=1) and
``DenseProjectionHead`` (heads.py:113-114, ``[B, C, H, W]``, dim=1).
"""
from __future__ import annotations

from torch import Tensor, nn

from . import ops

__all__ = ["Normalize", "normalize"]


def normalize(input: Tensor, p: float = 2.0, dim: int = 1, eps: float = 1e-12) -> Tensor:
    """``x / max(||x||_2, eps)`` along ``dim`` (same contract as ``torch.nn.functional.normalize``)."""
    if p != 2 and p != 2.0:
        raise NotImplementedError("only the L2 norm is on the hot path")
    y, _ = ops.l2norm_fwd(input, dim, eps)
    return y


class Normalize(nn.Module):
    def __init__(self, dim=1) -> None:
        super().__init__()
        self._dim = dim

    def forward(self, input):
        return normalize(input, p=2, dim=self._dim)
