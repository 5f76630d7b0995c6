"""fp64 restatement of the soft-positive-weight SupCon losses of ``contrastyou/losses/contrast_loss.py``.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) -- never imported by the product.

* ``SupConLoss2.forward`` .... :42-100   label / tri-state mask form, "in" (:90-92) and "out" (:94-97) modes
* ``SupConLoss3.forward`` .... :135-182  ``pos_weight`` [n, n] tiled 2 x 2 (:152), denominator over every j != i (:164)
* ``SupConLoss4.forward`` .... :206-262  weight blocks + ``enable_mask`` on the denominator (:246); the view-1 block
                                          is filled only when ``one2two_weight`` is given (:217-219, kept as it is)

All three reduce to: weights w_ij (diagonal removed), denominator mask en_ij (diagonal removed),
rowsum_i = sum_j en_ij exp(S_ij), W_i = sum_j w_ij,
  out: l_i = sum_j w_ij (S_ij - log rowsum_i) / W_i          in: l_i = log(sum_j w_ij exp(S_ij) / rowsum_i) / W_i
loss = -mean_i l_i; gradient dZ = (dS + dS^T) Z / tau with dS_ij = -(1/N) dl_i/dS_ij.
Pinned to outputs of the unmodified reference file by ``tests/golden/soft_weight_cases.npz``
(``oracle/make_golden_soft.py``).
"""
from __future__ import annotations

import numpy as np

__all__ = ["weighted_supcon", "weights_loss2", "weights_loss3", "weights_loss4"]


def weighted_supcon(z1, z2, w, en, *, temperature=0.07, in_mode=False, grad_out=1.0):
    """``w``, ``en``: [N, N] float64 / bool (their diagonals are ignored).  -> dict(loss, dz1, dz2)."""
    Z = np.concatenate([np.asarray(z1, np.float64), np.asarray(z2, np.float64)])
    N = Z.shape[0]
    off = ~np.eye(N, dtype=bool)
    w = np.asarray(w, np.float64) * off
    en = np.asarray(en, bool) & off
    S = Z @ Z.T / temperature
    m = S.max()
    E = np.exp(S - m)
    rowsum = (E * en).sum(1)
    W = w.sum(1)
    with np.errstate(divide="ignore", invalid="ignore"):
        if in_mode:
            q = (w * E).sum(1)
            l = np.log(q / rowsum) / W
            dl = (w * E / q[:, None] - en * E / rowsum[:, None]) / W[:, None]
        else:
            l = (w * (S - m - np.log(rowsum)[:, None])).sum(1) / W
            dl = w / W[:, None] - en * E / rowsum[:, None]
    loss = -l.mean()
    dS = -dl / N * grad_out
    dZ = (dS + dS.T) @ Z / temperature
    n = N // 2
    return dict(loss=float(loss), dz1=dZ[:n], dz2=dZ[n:])


def weights_loss2(n, target=None, mask=None):
    """(w, en) of SupConLoss2 (:55-77)."""
    if mask is not None:
        m = np.tile(np.asarray(mask), (2, 2))
        pos, neg = m == 1, m == 0
    elif target is not None:
        t = np.asarray(target, dtype=np.float32) if isinstance(target, (list, tuple)) else np.asarray(target)
        eq = np.tile(t[:, None] == t[None, :], (2, 2))
        pos, neg = eq, ~eq
    else:
        pos = np.tile(np.eye(n, dtype=bool), (2, 2))
        neg = ~pos
    return pos.astype(np.float64), pos | neg


def weights_loss3(pos_weight):
    """(w, en) of SupConLoss3 (:152, :164)."""
    w = np.tile(np.asarray(pos_weight, np.float64), (2, 2))
    return w, np.ones_like(w, dtype=bool)


def weights_loss4(n, one2one=None, two2two=None, one2two=None):
    """(w, en) of SupConLoss4 (:214-229), including its use of ``one2two_weight`` as the switch of the view-1 block."""
    w = np.zeros((2 * n, 2 * n))
    en = np.zeros((2 * n, 2 * n), dtype=bool)
    if one2two is not None:
        w[:n, :n] = one2one
        en[:n, :n] = True
    if two2two is not None:
        w[n:, n:] = two2two
        en[n:, n:] = True
    if one2two is not None:
        w[:n, n:] = one2two
        w[n:, :n] = one2two
        en[:n, n:] = True
        en[n:, :n] = True
    return w, en
