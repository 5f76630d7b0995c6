"""CPU restatement (numpy) of the dense-contrast front end.  TEST INFRASTRUCTURE -- only tests/, smoke() and
bench.py's cpu_baseline leg may import it; the product path never does.

Follows (paths under /root/reference):
  * contrastyou/projectors/heads.py:112      ``AdaptiveAvgPool2d(spatial_size)`` -- window of output cell i over
    an extent L split in n cells is [floor(i*L/n), ceil((i+1)*L/n)) (torch's adaptive pooling rule);
  * contrastyou/projectors/heads.py:113-114, nn.py:35-36   ``F.normalize(p=2, dim=1)``: x / max(||x||, 1e-12);
  * semi_seg/hooks/infonce.py:20-22, :233-241   point draw and gather (``region_extractor``);
  * contrastyou/epocher/comparable.py:398-404   all-pixels reshape ``[b,c,h,w] -> [b*h*w, c]``.
Pinned by tests/golden/dense_points.npz (made by oracle/make_golden_dense.py from the reference's own functions).
"""
from __future__ import annotations

import numpy as np


def _windows(L: int, n: int):
    return [((i * L) // n, -((-(i + 1) * L) // n)) for i in range(n)]


def adaptive_avg_pool(x: np.ndarray, ph: int, pw: int) -> np.ndarray:
    """x [B, C, H, W] -> [B, C, ph, pw] in float64."""
    B, C, H, W = x.shape
    x = x.astype(np.float64)
    out = np.empty((B, C, ph, pw))
    for i, (hs, he) in enumerate(_windows(H, ph)):
        for j, (ws, we) in enumerate(_windows(W, pw)):
            out[:, :, i, j] = x[:, :, hs:he, ws:we].mean(axis=(2, 3))
    return out


def dense_rows(x: np.ndarray, ph: int, pw: int, points=None, eps: float = 1e-12) -> np.ndarray:
    """Unit rows [B*P, C] (float64): every pooled pixel in (b, i, j) order, or ``points[b, p] = i*pw + j``."""
    pooled = adaptive_avg_pool(x, ph, pw)
    norm = np.maximum(np.sqrt((pooled ** 2).sum(axis=1, keepdims=True)), eps)
    y = (pooled / norm).transpose(0, 2, 3, 1).reshape(x.shape[0], ph * pw, x.shape[1])
    if points is None:
        return y.reshape(-1, x.shape[1])
    points = np.asarray(points)
    return np.stack([y[b, points[b]] for b in range(x.shape[0])]).reshape(-1, x.shape[1])


def dense_rows_grad(x: np.ndarray, ph: int, pw: int, gy: np.ndarray, points=None, eps: float = 1e-12) -> np.ndarray:
    """d<gy, dense_rows(x)>/dx in float64 (closed form: normalise backward, then the pooling adjoint)."""
    B, C, H, W = x.shape
    pooled = adaptive_avg_pool(x, ph, pw)                                   # [B, C, ph, pw]
    norm = np.sqrt((pooled ** 2).sum(axis=1, keepdims=True))
    inv = 1.0 / np.maximum(norm, eps)
    y = pooled * inv
    g_rows = np.zeros((B, ph * pw, C))
    if points is None:
        g_rows[:] = gy.reshape(B, ph * pw, C)
    else:
        points = np.asarray(points)
        gyr = gy.reshape(B, points.shape[1], C)
        for b in range(B):
            np.add.at(g_rows[b], points[b], gyr[b])
    g_y = g_rows.reshape(B, ph, pw, C).transpose(0, 3, 1, 2)
    g_pooled = inv * (g_y - y * (y * g_y).sum(axis=1, keepdims=True))
    gx = np.zeros((B, C, H, W))
    for i, (hs, he) in enumerate(_windows(H, ph)):
        for j, (ws, we) in enumerate(_windows(W, pw)):
            gx[:, :, hs:he, ws:we] += g_pooled[:, :, i, j][:, :, None, None] / ((he - hs) * (we - ws))
    return gx
