"""Blockwise fp64 torch reference of the loss + gradient (test helper).

Same closed form as oracle/closed_form.py (label masks only), but in torch so it can run on the GPU at
BASELINE's full sizes (N = 32768) in seconds.  It is itself checked against the numpy oracle in
test_gpu_parity.py::test_torch_ref_matches_oracle.
"""
import torch


def _weight(l, gamma, mode):
    if mode == "none":
        return torch.ones_like(l)
    if mode == "hard":
        return (l <= gamma).to(l.dtype)
    return torch.clamp_min(1 - l / gamma, 0)


@torch.no_grad()
def supcon_ref64(z1, z2, labels, *, temperature=0.07, gamma=1e6, mode="hard", correct_grad=False, block=2048,
                 want_grad=True):
    Z = torch.cat([z1, z2]).double()
    lab = torch.cat([labels, labels]).long()
    N = Z.shape[0]
    it = 1.0 / temperature
    idx = torch.arange(N, device=Z.device)
    logD = torch.empty(N, dtype=torch.float64, device=Z.device)
    c = torch.empty_like(logD); wl = torch.empty_like(logD); wp = torch.empty_like(logD)
    for i0 in range(0, N, block):
        i1 = min(N, i0 + block)
        S = Z[i0:i1] @ Z.t() * it
        off = idx[i0:i1, None] != idx[None, :]
        pos = (lab[i0:i1, None] == lab[None, :]) & off
        logD[i0:i1] = torch.logsumexp(S.masked_fill(~off, float("-inf")), dim=1)
        c[i0:i1] = pos.sum(1)
        llh = S - logD[i0:i1, None]
        W = _weight(-llh, gamma, mode)
        wl[i0:i1] = (W * llh * pos).sum(1)
        wp[i0:i1] = (W * pos).sum(1)
    ratio = wp.sum() / c.sum()
    scale = 1.0 / ratio if (correct_grad and ratio > 0) else 1.0
    loss = -(wl / c).mean() * scale
    out = dict(loss=loss.item(), ratio=ratio.item(), logD=logD, c=c)
    if not want_grad:
        return out
    A = wp / c
    k = float(scale) / N
    dZ = torch.empty_like(Z)
    for i0 in range(0, N, block):
        i1 = min(N, i0 + block)
        S = Z[i0:i1] @ Z.t() * it
        off = idx[i0:i1, None] != idx[None, :]
        pos = (lab[i0:i1, None] == lab[None, :]) & off
        llh_r = S - logD[i0:i1, None]
        llh_c = S - logD[None, :]
        T = (off * (A[i0:i1, None] * torch.exp(llh_r) + A[None, :] * torch.exp(llh_c))
             - pos * (_weight(-llh_r, gamma, mode) / c[i0:i1, None] + _weight(-llh_c, gamma, mode) / c[None, :]))
        dZ[i0:i1] = (T @ Z) * (k * it)
    n = N // 2
    out["dz1"], out["dz2"] = dZ[:n], dZ[n:]
    return out


@torch.no_grad()
def supcon_ref64_rows(Z, lab, row_begin, row_end, *, temperature=0.07, gamma=1e6, mode="hard", correct_grad=False,
                      block=2048, col_block=32768, reduce_sum=None, gather_rows=None):
    """The same fp64 closed form for the anchor rows [row_begin, row_end) of a (possibly sharded) problem.

    ``Z`` [N, d] / ``lab`` [N] hold ALL anchors in the global order (every rank has them after the all-gather).
    ``reduce_sum(t)`` sums a small fp64 tensor over the ranks and ``gather_rows(t)`` concatenates per-rank
    ``[k, rows]`` statistics along the row axis in rank order (both default to the single-process identity).
    Column blocks bound the temporaries at ``block x col_block`` doubles, so N = 262144 fits next to the product's
    own buffers.  -> dict(loss, ratio, logD[rows], dZ[rows, d])."""
    Z = Z.double()
    lab = lab.long()
    N = Z.shape[0]
    it = 1.0 / temperature
    dev = Z.device
    rows = row_end - row_begin
    reduce_sum = reduce_sum or (lambda t: t)
    gather_rows = gather_rows or (lambda t: t)
    logD = torch.empty(rows, dtype=torch.float64, device=dev)
    c = torch.zeros_like(logD); wl = torch.zeros_like(logD); wp = torch.zeros_like(logD)
    cols = torch.arange(N, device=dev)
    # pass 1: logD (streaming logsumexp over column blocks)
    for i0 in range(row_begin, row_end, block):
        i1 = min(row_end, i0 + block)
        m = torch.full((i1 - i0,), float("-inf"), dtype=torch.float64, device=dev)
        s = torch.zeros(i1 - i0, dtype=torch.float64, device=dev)
        for j0 in range(0, N, col_block):
            j1 = min(N, j0 + col_block)
            S = Z[i0:i1] @ Z[j0:j1].t() * it
            S.masked_fill_(cols[i0:i1, None] == cols[None, j0:j1], float("-inf"))
            mb = torch.maximum(m, S.max(dim=1).values)
            s = s * torch.exp(m - mb) + torch.exp(S - mb[:, None]).sum(1)
            m = mb
        logD[i0 - row_begin:i1 - row_begin] = m + torch.log(s)
    # pass 2: positives (needs the final logD)
    for i0 in range(row_begin, row_end, block):
        i1 = min(row_end, i0 + block)
        r = slice(i0 - row_begin, i1 - row_begin)
        for j0 in range(0, N, col_block):
            j1 = min(N, j0 + col_block)
            pos = (lab[i0:i1, None] == lab[None, j0:j1]) & (cols[i0:i1, None] != cols[None, j0:j1])
            if not bool(pos.any()):
                continue
            llh = Z[i0:i1] @ Z[j0:j1].t() * it - logD[r, None]
            W = _weight(-llh, gamma, mode)
            c[r] += pos.sum(1)
            wl[r] += (W * llh * pos).sum(1)
            wp[r] += (W * pos).sum(1)
    sums = reduce_sum(torch.stack([(wl / c).sum(), wp.sum(), c.sum()]))
    ratio = (sums[1] / sums[2]).item()
    scale = 1.0 / ratio if (correct_grad and ratio > 0) else 1.0
    loss = -(sums[0] / N).item() * scale
    stats_all = gather_rows(torch.stack([logD, c, wp / c]))          # [3, N] in the global row order
    logD_all, c_all, A_all = stats_all[0], stats_all[1], stats_all[2]
    k = float(scale) / N
    dZ = torch.zeros(rows, Z.shape[1], dtype=torch.float64, device=dev)
    A_r, c_r = wp / c, c
    for i0 in range(row_begin, row_end, block):
        i1 = min(row_end, i0 + block)
        r = slice(i0 - row_begin, i1 - row_begin)
        for j0 in range(0, N, col_block):
            j1 = min(N, j0 + col_block)
            S = Z[i0:i1] @ Z[j0:j1].t() * it
            off = cols[i0:i1, None] != cols[None, j0:j1]
            llh_r = S - logD[r, None]
            llh_c = S - logD_all[None, j0:j1]
            T = off * (A_r[r, None] * torch.exp(llh_r) + A_all[None, j0:j1] * torch.exp(llh_c))
            pos = (lab[i0:i1, None] == lab[None, j0:j1]) & off
            if bool(pos.any()):
                T -= pos * (_weight(-llh_r, gamma, mode) / c_r[r, None] + _weight(-llh_c, gamma, mode) / c_all[None, j0:j1])
            dZ[r] += (T @ Z[j0:j1]) * (k * it)
    return dict(loss=loss, ratio=ratio, logD=logD, dZ=dZ)
