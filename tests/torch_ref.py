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
