"""fp64 blockwise closed form of the self-paced SupCon loss and its gradient.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) -- never imported by the product.

Restates, without building any N x N array for large N, what the reference
computes in ``contrastyou/losses/contrast_loss3.py``:

* mask construction ............ ``SelfPacedSupConLoss.forward``   :126-145
* tiling + diagonal exclusion .. ``SelfPacedSupConLoss._forward``  :157-167
* similarity / temperature ..... ``exp_sim_temperature``           :25-31
* row sums, log-likelihood ..... ``_forward``                      :180-184
* self-paced weight ............ ``_self_paced_mask``              :207-214
* ratio, masked mean, scaling .. ``_forward``                      :189-201
* ``SupConLoss1`` (W == 1) ..... ``SupConLoss1._forward``          :61-110

Deviation from the reference bits (documented in SURVEY.md section 3.2): the
``+1e-16`` inside the log after the global-max shift (:184) is dropped; it is
< 1e-5 relative for tau >= 0.05.

Notation: Z = [z1; z2] (N = 2n rows), S = Z Z^T / tau,
P = positive mask, Q = negative mask, M = P | Q (both with the diagonal removed),
logD_i = logsumexp_{j in M_i} S_ij, LLH_ij = S_ij - logD_i, l_ij = -LLH_ij,
W_ij = hard 1[l_ij <= gamma] | soft max(0, 1 - l_ij / gamma) | none 1,
c_i = sum_j P_ij, loss = -(1/N) sum_i (1/c_i) sum_j W_ij P_ij LLH_ij,
ratio = sum W P / sum P, scale = 1/ratio if correct_grad and ratio > 0.
"""
from __future__ import annotations

import numpy as np

__all__ = ["codes_from_target", "supcon_closed_form", "self_paced_weight"]


def codes_from_target(target) -> np.ndarray:
    """Integer equality classes with the reference's label semantics.

    The reference turns a python list into ``torch.Tensor(list)`` (float32,
    contrast_loss3.py:134-135) and compares with ``torch.eq`` (:136); tensors are
    compared in their own dtype.  Two anchors are positives iff those values
    compare equal, so any labelling is equivalent to its equality classes.
    """
    if isinstance(target, (list, tuple)):
        arr = np.asarray(target, dtype=np.float32)
    else:
        arr = np.asarray(target)
    if arr.ndim != 1:
        raise ValueError("target must be 1-D")
    _, inv = np.unique(arr, return_inverse=True)
    return inv.astype(np.int64)


def self_paced_weight(llh: np.ndarray, gamma: float, mode: str, pos: np.ndarray) -> np.ndarray:
    """contrast_loss3.py:207-214 -- W = max(w, 1 - P)."""
    l = -llh
    if mode == "none":
        w = np.ones_like(l)
    elif mode == "hard":
        w = (l <= gamma).astype(l.dtype)
    else:  # every non-"hard" string is soft in the reference (:210-213)
        w = np.maximum(1.0 - l / gamma, 0.0)
    return np.maximum(w, 1.0 - pos)


class _Masks:
    """Row-block views of P (positives) and M (valid = P|Q) incl. diagonal removal."""

    def __init__(self, n: int, codes=None, tri=None):
        self.n, self.N = n, 2 * n
        self.codes = None if codes is None else np.concatenate([codes, codes])
        self.tri = tri  # [n, n] tri-state: 1 pos / 0 neg / other ignored (:128-131)

    def rows(self, i0: int, i1: int):
        """pos[b, N], valid[b, N] for anchors i0..i1."""
        idx = np.arange(i0, i1)
        cols = np.arange(self.N)
        offdiag = idx[:, None] != cols[None, :]
        if self.tri is not None:
            sub = self.tri[idx % self.n][:, cols % self.n]
            pos = (sub == 1) & offdiag
            valid = ((sub == 1) | (sub == 0)) & offdiag
        else:
            pos = (self.codes[idx][:, None] == self.codes[None, :]) & offdiag
            valid = offdiag
        return pos, valid

    def cols_T(self, i0: int, i1: int):
        """posT[b, N] = P[j, i], validT[b, N] = M[j, i] for i in block, all j."""
        if self.tri is None:
            return self.rows(i0, i1)  # label masks are symmetric
        idx = np.arange(i0, i1)
        cols = np.arange(self.N)
        offdiag = idx[:, None] != cols[None, :]
        sub = self.tri[cols % self.n][:, idx % self.n].T
        pos = (sub == 1) & offdiag
        valid = ((sub == 1) | (sub == 0)) & offdiag
        return pos, valid


def supcon_closed_form(z1, z2, *, target=None, mask=None, temperature: float = 0.07,
                       gamma: float = 1e6, mode: str = "hard", correct_grad: bool = False,
                       block: int = 512, grad_out: float = 1.0, want_grad: bool = True,
                       row_range=None, anchor_labels=None) -> dict:
    """Loss, ratio and gradient in fp64.

    ``mode``: "none" (SupConLoss1), "excl" (SupConLoss1 with ``exclude_other_pos=True``, :97-100),
    "hard", anything else = soft.
    ``row_range``: optional (r0, r1) -- only these anchor rows contribute to the
    *partial* sums returned under ``partial`` (used by the row-sharding tests);
    ``loss``/``ratio`` are always the full-problem values.
    ``anchor_labels``: optional integer label per ANCHOR (length N = 2n) instead of per sample; used by
    the row-sharding tests, where the global anchor order is (rank, view, sample) rather than (view, sample).
    Precedence mask > target > identity follows contrast_loss3.py:128-143.
    """
    z1 = np.asarray(z1, dtype=np.float64)
    z2 = np.asarray(z2, dtype=np.float64)
    assert z1.shape == z2.shape and z1.ndim == 2
    n, d = z1.shape
    N = 2 * n
    Z = np.concatenate([z1, z2], axis=0)
    inv_tau = 1.0 / float(temperature)

    if anchor_labels is not None:
        masks = _Masks(n, codes=None)
        masks.codes = codes_from_target(np.asarray(anchor_labels))
        assert masks.codes.shape[0] == N
    elif mask is not None:
        tri = np.asarray(mask)
        assert tri.shape == (n, n)
        masks = _Masks(n, tri=tri)
    elif target is not None:
        masks = _Masks(n, codes=codes_from_target(target))
    else:
        masks = _Masks(n, codes=np.arange(n, dtype=np.int64))  # SimCLR (:140-143)

    if mode == "excl":
        return _excl_closed_form(Z, n, masks, inv_tau, block, grad_out, want_grad)

    logD = np.empty(N)
    c = np.empty(N)
    wl = np.empty(N)   # sum_j W P LLH
    wp = np.empty(N)   # sum_j W P
    blocks = [(i0, min(i0 + block, N)) for i0 in range(0, N, block)]

    # sweep A: logD_i and c_i
    for i0, i1 in blocks:
        S = (Z[i0:i1] @ Z.T) * inv_tau
        pos, valid = masks.rows(i0, i1)
        Sm = np.where(valid, S, -np.inf)
        mx = Sm.max(axis=1)
        mx = np.where(np.isfinite(mx), mx, 0.0)
        with np.errstate(divide="ignore"):
            logD[i0:i1] = mx + np.log(np.exp(Sm - mx[:, None]).sum(axis=1))
        c[i0:i1] = pos.sum(axis=1)

    # sweep B: self-paced sums
    for i0, i1 in blocks:
        S = (Z[i0:i1] @ Z.T) * inv_tau
        pos, _ = masks.rows(i0, i1)
        with np.errstate(invalid="ignore"):
            llh = S - logD[i0:i1, None]
            W = self_paced_weight(llh, gamma, mode, pos.astype(np.float64))
        wl[i0:i1] = np.where(pos, W * llh, 0.0).sum(axis=1)
        wp[i0:i1] = np.where(pos, W, 0.0).sum(axis=1)

    with np.errstate(invalid="ignore", divide="ignore"):
        loss = -np.mean(wl / c)
        ratio = wp.sum() / c.sum()
    scale = 1.0
    if correct_grad and ratio > 0:
        scale = 1.0 / ratio
    loss = loss * scale

    out = dict(loss=float(loss), ratio=float(ratio), logD=logD, c=c, wl=wl, wp=wp, scale=scale)
    if row_range is not None:
        r0, r1 = row_range
        with np.errstate(invalid="ignore", divide="ignore"):
            out["partial"] = dict(loss_sum=float((wl[r0:r1] / c[r0:r1]).sum()),
                                  wp_sum=float(wp[r0:r1].sum()), p_sum=float(c[r0:r1].sum()))
    if not want_grad:
        return out

    # sweep C: dZ_i = (1/tau) sum_j (dS_ij + dS_ji) z_j
    #   dS_ij = -(g*scale/N) [ W_ij P_ij / c_i - A_i M_ij exp(S_ij - logD_i) ],  A_i = wp_i / c_i
    with np.errstate(invalid="ignore", divide="ignore"):
        A = wp / c
    k = grad_out * scale / N
    dZ = np.empty_like(Z)
    for i0, i1 in blocks:
        S = (Z[i0:i1] @ Z.T) * inv_tau
        pos, valid = masks.rows(i0, i1)
        posT, validT = masks.cols_T(i0, i1)
        with np.errstate(invalid="ignore", over="ignore"):
            llh_r = S - logD[i0:i1, None]            # LLH_ij
            llh_c = S - logD[None, :]                # LLH_ji (S symmetric)
            W_r = self_paced_weight(llh_r, gamma, mode, pos.astype(np.float64))
            W_c = self_paced_weight(llh_c, gamma, mode, posT.astype(np.float64))
            dS_r = -k * (np.where(pos, W_r / c[i0:i1, None], 0.0)
                         - np.where(valid, A[i0:i1, None] * np.exp(llh_r), 0.0))
            dS_c = -k * (np.where(posT, W_c / c[None, :], 0.0)
                         - np.where(validT, A[None, :] * np.exp(llh_c), 0.0))
        dZ[i0:i1] = ((dS_r + dS_c) @ Z) * inv_tau
    out["dz1"] = dZ[:n]
    out["dz2"] = dZ[n:]
    return out


def _excl_closed_form(Z, n, masks, inv_tau, block, grad_out, want_grad) -> dict:
    """``SupConLoss1(exclude_other_pos=True)`` -- contrast_loss3.py:93-106.

    E = exp(S - shift) (any common shift cancels), negsum_i = sum_{j in Q_i} E_ij,
    r_i = q_i / (c_i + q_i) (:98), B_i = negsum_i / (r_i + 1e-4) (:100),
    LLH_ij = (S_ij - shift) - log(E_ij + B_i) (:99-100, the +1e-16 dropped as in the default mode),
    loss = -(1/N) sum_i (1/c_i) sum_{j in P_i} LLH_ij (:105-106).
    dLoss/dS_ij = -(g/N) [ P_ij B_i / (c_i (E_ij + B_i)) - Q_ij E_ij v_i ],
    v_i = (1/c_i) sum_{j in P_i} 1 / (E_ij + B_i) / (r_i + 1e-4).
    """
    N = 2 * n
    shift = inv_tau
    blocks = [(i0, min(i0 + block, N)) for i0 in range(0, N, block)]
    c = np.empty(N)
    q = np.empty(N)
    negsum = np.empty(N)
    for i0, i1 in blocks:
        E = np.exp((Z[i0:i1] @ Z.T) * inv_tau - shift)
        pos, valid = masks.rows(i0, i1)
        neg = valid & ~pos
        c[i0:i1] = pos.sum(axis=1)
        q[i0:i1] = neg.sum(axis=1)
        negsum[i0:i1] = np.where(neg, E, 0.0).sum(axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        rr = q / (c + q) + 1e-4
        B = negsum / rr
    wl = np.empty(N)
    v = np.empty(N)
    for i0, i1 in blocks:
        S = (Z[i0:i1] @ Z.T) * inv_tau - shift
        pos, _ = masks.rows(i0, i1)
        with np.errstate(invalid="ignore", divide="ignore"):
            den = np.exp(S) + B[i0:i1, None]
            wl[i0:i1] = np.where(pos, S - np.log(den), 0.0).sum(axis=1)
            v[i0:i1] = np.where(pos, 1.0 / den, 0.0).sum(axis=1) / (c[i0:i1] * rr[i0:i1])
    with np.errstate(invalid="ignore", divide="ignore"):
        loss = -np.mean(wl / c)
    out = dict(loss=float(loss), ratio=1.0, c=c, q=q, negsum=negsum, B=B, v=v, wl=wl, wp=c.copy(), scale=1.0)
    if not want_grad:
        return out
    k = grad_out / N
    dZ = np.empty_like(Z)
    for i0, i1 in blocks:
        E = np.exp((Z[i0:i1] @ Z.T) * inv_tau - shift)
        pos, valid = masks.rows(i0, i1)
        posT, validT = masks.cols_T(i0, i1)
        neg, negT = valid & ~pos, validT & ~posT
        with np.errstate(invalid="ignore", divide="ignore"):
            Bi, Bj = B[i0:i1, None], B[None, :]
            dS_r = -k * (np.where(pos, Bi / (c[i0:i1, None] * (E + Bi)), 0.0) - np.where(neg, E * v[i0:i1, None], 0.0))
            dS_c = -k * (np.where(posT, Bj / (c[None, :] * (E + Bj)), 0.0) - np.where(negT, E * v[None, :], 0.0))
        dZ[i0:i1] = ((dS_r + dS_c) @ Z) * inv_tau
    out["dz1"] = dZ[:n]
    out["dz2"] = dZ[n:]
    return out
