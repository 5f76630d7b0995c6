"""torch custom ops over the C ABI (``include/spcl.h``).

``spcl::supcon_fwd`` / ``spcl::supcon_bwd`` are the fused self-paced SupCon forward / backward
(reference: ``contrastyou/losses/contrast_loss3.py:25-31, :147-214`` and their autograd),
``spcl::l2norm_fwd`` / ``spcl::l2norm_bwd`` the projector's normalise tail
(``contrastyou/projectors/nn.py:35-36``).  PyTorch is used for device memory, streams and autograd
plumbing only; all arithmetic on the path happens in ``libspcl_b200.so``.
"""
from __future__ import annotations

import ctypes
import functools
import math
import os
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _native as nat

__all__ = ["supcon_fwd", "supcon_bwd", "l2norm_fwd", "l2norm_bwd", "pad_to", "label_codes", "tri_codes"]


def _ptr(t: Optional[Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(t: Tensor):
    # (the raw-handle query: torch.cuda.current_stream() builds a Stream object, ~10 us per call on the host)
    idx = t.device.index
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device() if idx is None else idx))


def pad_to(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def on_device_of(idx: int):
    """Runs the wrapped op with the CUDA device of its ``idx``-th positional tensor current.

    The C ABI launches kernels, memsets, ``cudaFuncSetAttribute`` and tensor-map encodes on the CURRENT device;
    tensors that live on another GPU of the process (``model.to('cuda:1')`` without ``set_device``, one module per
    device) would otherwise hit an invalid-resource-handle error where the reference (pure torch) just works."""
    def deco(fn):
        @functools.wraps(fn)
        def wrapped(*args, **kwargs):
            t = args[idx]
            if not (isinstance(t, Tensor) and t.is_cuda) or t.device.index == torch.cuda.current_device():
                return fn(*args, **kwargs)
            with torch.cuda.device(t.device):
                return fn(*args, **kwargs)
        return wrapped
    return deco


def _require_cuda(*ts: Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("spcl_b200 ops run on CUDA tensors only (there is no CPU fallback); got a "
                               f"{t.device} tensor")


# --------------------------------------------------------------------------------------------------
# label handling (host side of contrast_loss3.py:133-139)
# --------------------------------------------------------------------------------------------------
def label_codes(target, n: int, device) -> Tensor:
    """int32[n] codes whose equality classes equal the reference's ``torch.eq`` on ``target``.

    python lists go through float32 in the reference (``torch.Tensor(list)``, :135); integers below
    2**24 are exact there, so they are used as they are; anything else is mapped to its float32
    equality classes.  Tensors are compared in their own dtype (:136).
    """
    import numpy as np

    if isinstance(target, (list, tuple)):
        arr = np.asarray(target)
        if arr.ndim != 1 or arr.shape[0] != n:
            raise AssertionError((arr.shape, n))
        if arr.dtype.kind in "iub" and (arr.size == 0 or np.abs(arr).max() < 2 ** 24):
            codes = arr.astype(np.int32)
        else:
            _, inv = np.unique(arr.astype(np.float32), return_inverse=True)
            codes = inv.astype(np.int32)
        return torch.from_numpy(codes).to(device, non_blocking=True)
    if not isinstance(target, Tensor):
        raise TypeError(f"target must be a list or a Tensor, got {type(target)}")
    if target.dim() != 1 or target.shape[0] != n:
        raise AssertionError((tuple(target.shape), n))
    t = target.to(device)
    if t.dtype in (torch.int32, torch.int16, torch.int8, torch.uint8, torch.bool):
        return t.to(torch.int32)
    if t.dtype == torch.int64 and (t.numel() == 0 or int(t.abs().max()) < 2 ** 31 - 1):
        # torch's default label dtype: a plain cast keeps the equality classes (one scalar read, no sort);
        # pass int32 labels to avoid even that synchronisation
        return t.to(torch.int32)
    # wide int64 / floating labels: exact equality classes through a sort (this path synchronises)
    _, inv = torch.unique(t, return_inverse=True)
    return inv.to(torch.int32)


def tri_codes(mask: Tensor, n: int, device) -> Tensor:
    """uint8[n, n]: 1 = positive (mask == 1), 0 = negative (mask == 0), 2 = ignored (:130-131)."""
    if tuple(mask.shape) != (n, n):
        raise AssertionError((tuple(mask.shape), n))
    m = mask.to(device)
    out = torch.full((n, n), 2, dtype=torch.uint8, device=device)
    out[m == 0] = 0
    out[m == 1] = 1
    return out


# --------------------------------------------------------------------------------------------------
# fused loss
# --------------------------------------------------------------------------------------------------
def _f32_split(has_labels: bool, mode: int) -> bool:
    """fp32 path: label form with W in {1, hard, soft} runs on the 2-D grid kernels; the tri-state ``mask=`` form and
    ``exclude_other_pos`` keep the row-grid kernels.  SPCL_F32_SPLIT=0 forces the row grid (A/B switch)."""
    return has_labels and mode != nat.MODE_EXCL and os.environ.get("SPCL_F32_SPLIT", "1") != "0"



@torch.library.custom_op("spcl::supcon_fwd", mutates_args=(), device_types="cuda")
@on_device_of(0)
def supcon_fwd(z1: Tensor, z2: Tensor, labels: Optional[Tensor], tri: Optional[Tensor], temperature: float,
               gamma: float, mode: int, correct_grad: bool, use_tc: bool
               ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """-> (scalars[4] = loss, ratio, scale, scale/N ; row_stats[4, n_pad] ; packed operands ;
    labels_full int32[n_pad] ; block signatures int32[n_pad/128, 4])."""
    _require_cuda(z1, z2, labels, tri)
    if z1.shape != z2.shape or z1.dim() != 2:
        raise AssertionError((tuple(z1.shape), tuple(z2.shape)))
    if z1.dtype != torch.float32 or z2.dtype != torch.float32:
        raise TypeError("spcl::supcon_fwd takes float32 anchors")
    n, d = z1.shape
    N = 2 * n
    if d > nat.MAX_D:
        raise nat.SpclError(f"embedding width {d} > {nat.MAX_D} is not supported by this build")
    if (labels is None) == (tri is None):
        raise ValueError("exactly one of labels / tri must be given")
    if tri is not None and use_tc:
        raise nat.SpclError("the tri-state mask= form runs on the fp32 path only (precision='fp32' or 'auto')")
    dev = z1.device
    z1 = z1.contiguous()
    z2 = z2.contiguous()
    n_pad = pad_to(N, nat.TILE)
    st = _stream(z1)
    inv_tau = 1.0 / float(temperature)
    scalars = torch.empty(4, dtype=torch.float32, device=dev)
    if labels is not None and (labels.dtype != torch.int32 or labels.shape != (n,)):
        raise TypeError("labels must be int32[n]")

    if use_tc:
        # one launch packs both views to bf16, tiles the labels, builds the block signatures and zeroes the
        # partial sums; nothing on this path is initialised by a torch fill kernel
        d_pad = pad_to(d, 64)
        row_stats = torch.empty(4, n_pad, dtype=torch.float32, device=dev)
        partials = torch.empty(3, dtype=torch.float32, device=dev)
        zpack = torch.empty(n_pad, d_pad, dtype=torch.bfloat16, device=dev)
        labels_full = torch.empty(n_pad, dtype=torch.int32, device=dev)
        sig = torch.empty(n_pad // nat.TILE, 4, dtype=torch.int32, device=dev)
        acc = torch.empty(n_pad, 4, dtype=torch.float32, device=dev)
        nat.call("spcl_supcon_prepare_bf16", _ptr(z1), _ptr(z2), n, d, z1.stride(0), z2.stride(0), _ptr(labels),
                 _ptr(zpack), n_pad, d_pad, _ptr(labels_full), _ptr(sig), _ptr(partials), st)
        nat.call("spcl_supcon_fwd_bf16", _ptr(zpack), N, n_pad, d_pad, _ptr(labels_full), _ptr(sig), 0, N,
                 inv_tau, float(gamma), int(mode), _ptr(acc), _ptr(row_stats), _ptr(partials), st)
    else:
        row_stats = torch.zeros(4, n_pad, dtype=torch.float32, device=dev)
        partials = torch.zeros(3, dtype=torch.float32, device=dev)
        if labels is not None:
            labels_full = torch.zeros(n_pad, dtype=torch.int32, device=dev)
            labels_full[:n] = labels
            labels_full[n:N] = labels
        else:
            labels_full = torch.empty(0, dtype=torch.int32, device=dev)
        zpack = torch.cat([z1, z2], dim=0)
        sig = torch.empty(0, 4, dtype=torch.int32, device=dev)
        if _f32_split(labels is not None, mode):
            # (row block, column range) grid: the batch sizes the reference trains with fill the machine
            acc = torch.empty(N, 4, dtype=torch.float32, device=dev)
            nat.call("spcl_supcon_fwd_f32_split", _ptr(zpack), N, d, zpack.stride(0), _ptr(labels_full), 0, N,
                     inv_tau, float(gamma), int(mode), _ptr(acc), _ptr(row_stats), n_pad, _ptr(partials), st)
        else:
            nat.call("spcl_supcon_fwd_f32", _ptr(zpack), N, d, zpack.stride(0),
                     _ptr(labels_full) if labels is not None else None, _ptr(tri), n, 0, N, inv_tau, float(gamma),
                     int(mode), _ptr(row_stats), n_pad, _ptr(partials), st)
    nat.call("spcl_supcon_finalize", _ptr(partials), N, int(bool(correct_grad)), _ptr(scalars), st)
    return scalars, row_stats, zpack, labels_full, sig


@supcon_fwd.register_fake
def _(z1, z2, labels, tri, temperature, gamma, mode, correct_grad, use_tc):
    n, d = z1.shape
    n_pad = pad_to(2 * n, nat.TILE)
    scalars = z1.new_empty(4)
    row_stats = z1.new_empty(4, n_pad)
    if use_tc:
        zpack = z1.new_empty(n_pad, pad_to(d, 64), dtype=torch.bfloat16)
        sig = z1.new_empty(n_pad // nat.TILE, 4, dtype=torch.int32)
    else:
        zpack = z1.new_empty(2 * n, d)
        sig = z1.new_empty(0, 4, dtype=torch.int32)
    labels_full = z1.new_empty(n_pad if labels is not None else 0, dtype=torch.int32)
    return scalars, row_stats, zpack, labels_full, sig


@torch.library.custom_op("spcl::supcon_bwd", mutates_args=(), device_types="cuda")
@on_device_of(1)
def supcon_bwd(grad_loss: Tensor, zpack: Tensor, labels_full: Tensor, sig: Tensor, tri: Optional[Tensor],
               row_stats: Tensor, scalars: Tensor, temperature: float, gamma: float, mode: int, use_tc: bool,
               n: int, d: int) -> Tensor:
    """-> dZ float32 [2n, d] (rows [0, n) belong to view 1, [n, 2n) to view 2)."""
    _require_cuda(grad_loss, zpack, row_stats, scalars)
    N = 2 * n
    dev = zpack.device
    st = _stream(zpack)
    inv_tau = 1.0 / float(temperature)
    g = grad_loss.reshape(1).to(torch.float32).contiguous()
    dz = torch.empty(N, d, dtype=torch.float32, device=dev)
    if use_tc:
        n_pad, d_pad = zpack.shape
        zt = torch.empty(nat.workspace_bytes(nat.WS_BWD_ZT, n_pad, d_pad), dtype=torch.uint8, device=dev)
        nat.call("spcl_supcon_bwd_bf16", _ptr(zpack), N, n_pad, d_pad, d, _ptr(labels_full), _ptr(sig),
                 _ptr(row_stats), _ptr(scalars), _ptr(g), 0, N, inv_tau, float(gamma), int(mode), _ptr(dz),
                 dz.stride(0), _ptr(zt), st)
    elif _f32_split(tri is None, mode):
        nat.call("spcl_supcon_bwd_f32_split", _ptr(zpack), N, d, zpack.stride(0), _ptr(labels_full), _ptr(row_stats),
                 row_stats.stride(0), _ptr(scalars), _ptr(g), 0, N, inv_tau, float(gamma), int(mode), _ptr(dz),
                 dz.stride(0), st)
    else:
        nat.call("spcl_supcon_bwd_f32", _ptr(zpack), N, d, zpack.stride(0),
                 _ptr(labels_full) if tri is None else None, _ptr(tri), n, _ptr(row_stats),
                 row_stats.stride(0), _ptr(scalars), _ptr(g), 0, N, inv_tau, float(gamma), int(mode), _ptr(dz), dz.stride(0), st)
    return dz


@supcon_bwd.register_fake
def _(grad_loss, zpack, labels_full, sig, tri, row_stats, scalars, temperature, gamma, mode, use_tc, n, d):
    return row_stats.new_empty(2 * n, d)


def _supcon_setup(ctx, inputs, output):
    z1, z2, labels, tri, temperature, gamma, mode, correct_grad, use_tc = inputs
    scalars, row_stats, zpack, labels_full, sig = output
    ctx.save_for_backward(zpack, labels_full, sig, tri, row_stats, scalars)
    ctx.hp = (temperature, gamma, mode, use_tc, z1.shape[0], z1.shape[1])
    ctx.set_materialize_grads(False)


def _supcon_backward(ctx, g_scalars, g_stats, g_zpack, g_labels, g_sig):
    if g_scalars is None:
        return (None,) * 9
    zpack, labels_full, sig, tri, row_stats, scalars = ctx.saved_tensors
    temperature, gamma, mode, use_tc, n, d = ctx.hp
    # only d loss / d z is defined; ratio / scale are constants of the step (the reference detaches them)
    dz = supcon_bwd(g_scalars[0], zpack, labels_full, sig, tri, row_stats, scalars, temperature, gamma, mode,
                    use_tc, n, d)
    return dz[:n], dz[n:], None, None, None, None, None, None, None


supcon_fwd.register_autograd(_supcon_backward, setup_context=_supcon_setup)


# --------------------------------------------------------------------------------------------------
# one cooperative launch for forward + backward at the reference's own batch sizes (spcl_supcon_group_fused_f32)
# --------------------------------------------------------------------------------------------------
_FUSED_CAP = {}


def fused_capacity(device) -> int:
    """CTAs of the fused small-batch kernel that can be co-resident on ``device`` (cached per device)."""
    idx = torch.device(device).index
    idx = torch.cuda.current_device() if idx is None else idx
    if idx not in _FUSED_CAP:
        with torch.cuda.device(idx):
            _FUSED_CAP[idx] = int(nat.lib().spcl_supcon_fused_capacity())
    return _FUSED_CAP[idx]


def fused_fits(shapes, device) -> bool:
    """True when the K problems of ``shapes`` = [(n, d), ...] (2n anchors each) run as ONE cooperative launch: every
    64 x 64 tile is a CTA and all must be resident.  SPCL_FUSED_SMALL=0 switches the route off (A/B, tests of the
    multi-launch path)."""
    if os.environ.get("SPCL_FUSED_SMALL", "1") == "0" or not shapes:
        return False
    tiles = max((2 * n + 63) // 64 for n, _ in shapes)
    return tiles * tiles * len(shapes) <= fused_capacity(device)


def _fused_launch(probs, count: int, st, device) -> bool:
    """``spcl_supcon_group_fused_f32``; False when the library declines (the cooperative grid cannot be resident on
    this context after all, e.g. an SM-partitioned one): the route is then switched off for the device and the caller
    runs the staged entry points.  Any other error raises."""
    rc = nat.lib().spcl_supcon_group_fused_f32(ctypes.byref(probs), count, st)
    if rc == nat.ERR_UNSUPPORTED:
        idx = torch.device(device).index
        _FUSED_CAP[torch.cuda.current_device() if idx is None else idx] = 0
        return False
    nat.check(rc, "spcl_supcon_group_fused_f32")
    return True


def _fused_problem(q, z, labels2, N, d, temperature, gamma, mode, correct_grad, ws, dz):
    """Fill one ``spcl_problem_f32`` from the workspace ``ws`` = acc [N,4] | row_stats [4,N] | partials [4] | scalars [4]."""
    q.z, q.n_total, q.d, q.ldz, q.labels = z.data_ptr(), N, d, z.stride(0), labels2.data_ptr()
    q.inv_tau, q.gamma, q.mode, q.correct_grad = 1.0 / float(temperature), float(gamma), int(mode), int(bool(correct_grad))
    base = ws.data_ptr()                               # (pointer arithmetic: a tensor slice costs ~3 us on the host)
    q.acc, q.row_stats, q.stats_stride = base, base + N * 16, N
    q.partials, q.scalars = base + N * 32, base + N * 32 + 16
    q.dz, q.lddz = dz.data_ptr(), dz.stride(0)


class _FusedDeclined(Exception):
    pass


class _FusedSupCon(torch.autograd.Function):
    """One fp32 label-form problem on the fused kernel: the forward call also produces dz for an upstream gradient of
    1, the backward is one multiply."""

    @staticmethod
    @on_device_of(1)
    def forward(ctx, z1, z2, labels, temperature, gamma, mode, correct_grad):
        n, d = z1.shape
        N = 2 * n
        z = torch.cat([z1, z2], dim=0)                                      # contrast_loss3.py:26
        lab2 = torch.cat([labels, labels])                                  # (Tensor.repeat: 24 us of host time)
        ws = torch.empty(N * 8 + 8, dtype=torch.float32, device=z.device)
        dz = torch.empty(N, d, dtype=torch.float32, device=z.device)
        probs = (nat.ProblemF32 * 1)()
        _fused_problem(probs[0], z, lab2, N, d, temperature, gamma, mode, correct_grad, ws, dz)
        if not _fused_launch(probs, 1, _stream(z), z.device):
            raise _FusedDeclined()
        ctx.dz, ctx.n = dz, n
        ctx.set_materialize_grads(False)
        row_stats = ws[N * 4:N * 8].view(4, N)
        ctx.mark_non_differentiable(row_stats)
        return ws[N * 8 + 4:N * 8 + 8], row_stats

    @staticmethod
    def backward(ctx, g_scalars, _g_stats):
        if g_scalars is None:
            return (None,) * 7
        dz = ctx.dz * g_scalars[0]
        return dz[:ctx.n], dz[ctx.n:], None, None, None, None, None


class _DirectSupCon(torch.autograd.Function):
    """The same two entry points without the ``torch.library`` dispatch, for eager callers.

    A Python custom op costs ~100 us of dispatcher / fake-tensor bookkeeping per call on the host -- more than the
    kernels of one N = 512 problem.  ``spcl::supcon_fwd`` / ``spcl::supcon_bwd`` stay the registered op surface
    (``torch.compile``, ``opcheck``, export); the ``nn.Module``s call this when they run eagerly."""

    @staticmethod
    def forward(ctx, z1, z2, labels, tri, temperature, gamma, mode, correct_grad, use_tc):
        scalars, row_stats, zpack, labels_full, sig = supcon_fwd._init_fn(z1, z2, labels, tri, temperature, gamma,
                                                                          mode, correct_grad, use_tc)
        ctx.save_for_backward(zpack, labels_full, sig, tri, row_stats, scalars)
        ctx.hp = (temperature, gamma, mode, use_tc, z1.shape[0], z1.shape[1])
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(row_stats)
        return scalars, row_stats

    @staticmethod
    def backward(ctx, g_scalars, _g_stats):
        if g_scalars is None:
            return (None,) * 9
        zpack, labels_full, sig, tri, row_stats, scalars = ctx.saved_tensors
        temperature, gamma, mode, use_tc, n, d = ctx.hp
        dz = supcon_bwd._init_fn(g_scalars[0], zpack, labels_full, sig, tri, row_stats, scalars, temperature, gamma,
                                 mode, use_tc, n, d)
        return dz[:n], dz[n:], None, None, None, None, None, None, None


def supcon_fwd_eager(z1, z2, labels, tri, temperature, gamma, mode, correct_grad, use_tc):
    """-> (scalars[4], row_stats): ``spcl::supcon_fwd`` under tracing / compilation, the direct route otherwise."""
    if torch.compiler.is_compiling():
        scalars, row_stats, _, _, _ = supcon_fwd(z1, z2, labels, tri, temperature, gamma, mode, correct_grad, use_tc)
        return scalars, row_stats
    if (not use_tc and labels is not None and mode != nat.MODE_EXCL and z1.dim() == 2 and z1.shape == z2.shape
            and z1.dtype == torch.float32 and z2.dtype == torch.float32 and z1.shape[1] <= nat.MAX_D
            and fused_fits([tuple(z1.shape)], z1.device)):
        _require_cuda(z1, z2, labels)
        if labels.dtype != torch.int32 or labels.shape != (z1.shape[0],):
            raise TypeError("labels must be int32[n]")
        try:
            return _FusedSupCon.apply(z1.contiguous(), z2.contiguous(), labels, temperature, gamma, mode, correct_grad)
        except _FusedDeclined:
            pass                                     # (fused_fits is False for this device from now on)
    return _DirectSupCon.apply(z1, z2, labels, tri, temperature, gamma, mode, correct_grad, use_tc)


class _RawSupCon(torch.autograd.Function):
    """Fused projector tail + loss: un-normalised ``[B, C]`` / ``[B, C, H, W]`` projector outputs in,
    loss out; F.normalize, the NCHW -> [B*HW, C] reshape copy, torch.cat and the bf16 pack are one kernel
    (``spcl_supcon_prepare_raw_bf16``) and the normalise backward consumes the loss-gradient rows directly
    (``spcl_supcon_raw_bwd``).  Tensor-core path only."""

    @staticmethod
    @on_device_of(1)
    def forward(ctx, x1, x2, labels, temperature, gamma, mode, correct_grad, eps):
        _require_cuda(x1, x2, labels)
        if x1.shape != x2.shape or x1.dim() < 2:
            raise AssertionError((tuple(x1.shape), tuple(x2.shape)))
        if x1.dtype != torch.float32 or x2.dtype != torch.float32:
            raise TypeError("the fused projector tail takes float32 projector outputs")
        x1, x2 = x1.contiguous(), x2.contiguous()
        outer, d = x1.shape[0], x1.shape[1]
        inner = math.prod(x1.shape[2:])
        n = outer * inner
        N = 2 * n
        if d > nat.MAX_D:
            raise nat.SpclError(f"embedding width {d} > {nat.MAX_D} is not supported by this build")
        if labels is not None and (labels.dtype != torch.int32 or labels.shape != (n,)):
            raise TypeError("labels must be int32[outer * inner] (one per anchor)")
        dev, st = x1.device, _stream(x1)
        n_pad, d_pad = pad_to(N, nat.TILE), pad_to(d, 64)
        inv_tau = 1.0 / float(temperature)
        zpack = torch.empty(n_pad, d_pad, dtype=torch.bfloat16, device=dev)
        labels_full = torch.empty(n_pad, dtype=torch.int32, device=dev)
        sig = torch.empty(n_pad // nat.TILE, 4, dtype=torch.int32, device=dev)
        acc = torch.empty(n_pad, 4, dtype=torch.float32, device=dev)
        row_stats = torch.empty(4, n_pad, dtype=torch.float32, device=dev)
        partials = torch.empty(3, dtype=torch.float32, device=dev)
        inv_norm = torch.empty(N, dtype=torch.float32, device=dev)
        scalars = torch.empty(4, dtype=torch.float32, device=dev)
        nat.call("spcl_supcon_prepare_raw_bf16", _ptr(x1), _ptr(x2), outer, d, inner, float(eps), _ptr(labels),
                 _ptr(zpack), n_pad, d_pad, _ptr(inv_norm), _ptr(labels_full), _ptr(sig), _ptr(partials), st)
        nat.call("spcl_supcon_fwd_bf16", _ptr(zpack), N, n_pad, d_pad, _ptr(labels_full), _ptr(sig), 0, N, inv_tau,
                 float(gamma), int(mode), _ptr(acc), _ptr(row_stats), _ptr(partials), st)
        nat.call("spcl_supcon_finalize", _ptr(partials), N, int(bool(correct_grad)), _ptr(scalars), st)
        ctx.save_for_backward(x1, x2, inv_norm, zpack, labels_full, sig, row_stats, scalars)
        ctx.hp = (float(temperature), float(gamma), int(mode), n, d, outer, inner)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(row_stats)
        return scalars, row_stats

    @staticmethod
    def backward(ctx, g_scalars, _g_stats):
        if g_scalars is None:
            return (None,) * 8
        x1, x2, inv_norm, zpack, labels_full, sig, row_stats, scalars = ctx.saved_tensors
        temperature, gamma, mode, n, d, outer, inner = ctx.hp
        dz = supcon_bwd._init_fn(g_scalars[0], zpack, labels_full, sig, None, row_stats, scalars, temperature, gamma,
                                 mode, True, n, d)
        gx1, gx2 = torch.empty_like(x1), torch.empty_like(x2)
        nat.call("spcl_supcon_raw_bwd", _ptr(dz), dz.stride(0), _ptr(x1), _ptr(x2), _ptr(inv_norm), _ptr(gx1),
                 _ptr(gx2), outer, d, inner, _stream(x1))
        return gx1, gx2, None, None, None, None, None, None


def supcon_fwd_raw(x1, x2, labels, temperature, gamma, mode, correct_grad, eps=1e-12):
    """-> (scalars[4], row_stats) from UN-normalised projector outputs ``[B, C]`` or ``[B, C, *spatial]``."""
    return _RawSupCon.apply(x1, x2, labels, temperature, gamma, mode, correct_grad, eps)


class _WeightedSupCon(torch.autograd.Function):
    """Soft positive weights (``SupConLoss3`` / ``SupConLoss4`` / ``SupConLoss2`` in-mode,
    contrastyou/losses/contrast_loss.py:33-270) on the fp32 kernels.  ``pw``: float [pwn, pwn] with pwn = n
    (tiled 2 x 2 like ``pos_weight.repeat(2, 2)``) or pwn = 2n; ``enable``: uint8 [2n, 2n] or None."""

    @staticmethod
    @on_device_of(1)
    def forward(ctx, z1, z2, pw, enable, temperature, in_mode):
        _require_cuda(z1, z2, pw, enable)
        if z1.shape != z2.shape or z1.dim() != 2:
            raise AssertionError((tuple(z1.shape), tuple(z2.shape)))
        n, d = z1.shape
        N = 2 * n
        if d > nat.MAX_D:
            raise nat.SpclError(f"embedding width {d} > {nat.MAX_D} is not supported by this build")
        pw = pw.to(torch.float32).contiguous()
        if pw.dim() != 2 or pw.shape[0] != pw.shape[1] or pw.shape[0] not in (n, N):
            raise AssertionError(tuple(pw.shape))
        if enable is not None:
            if tuple(enable.shape) != (N, N) or enable.dtype != torch.uint8:
                raise TypeError("enable must be uint8 [2n, 2n]")
            enable = enable.contiguous()
        dev, st = z1.device, _stream(z1)
        z = torch.cat([z1.float(), z2.float()], dim=0)               # contrast_loss.py:26
        row_stats = torch.empty(4, N, dtype=torch.float32, device=dev)
        partials = torch.zeros(4, dtype=torch.float32, device=dev)
        scalars = torch.empty(4, dtype=torch.float32, device=dev)
        inv_tau = 1.0 / float(temperature)
        nat.call("spcl_supcon_fwd_w_f32", _ptr(z), N, d, z.stride(0), _ptr(pw), pw.shape[0], _ptr(enable), int(in_mode),
                 inv_tau, _ptr(row_stats), N, _ptr(partials), st)
        nat.call("spcl_supcon_finalize", _ptr(partials), N, 0, _ptr(scalars), st)
        ctx.save_for_backward(z, pw, enable, row_stats, scalars)
        ctx.hp = (inv_tau, int(in_mode), n, d)
        return scalars[0].clone()

    @staticmethod
    def backward(ctx, g):
        z, pw, enable, row_stats, scalars = ctx.saved_tensors
        inv_tau, in_mode, n, d = ctx.hp
        N = 2 * n
        g = g.reshape(1).to(torch.float32).contiguous()
        dz = torch.empty(N, d, dtype=torch.float32, device=z.device)
        nat.call("spcl_supcon_bwd_w_f32", _ptr(z), N, d, z.stride(0), _ptr(pw), pw.shape[0], _ptr(enable), in_mode,
                 inv_tau, _ptr(row_stats), N, _ptr(scalars), _ptr(g), _ptr(dz), d, _stream(z))
        return dz[:n], dz[n:], None, None, None, None


def supcon_weighted(z1: Tensor, z2: Tensor, pw: Tensor, enable: Optional[Tensor], temperature: float, in_mode: bool):
    """-> 0-d loss of the soft-positive-weight SupCon family (differentiable w.r.t. z1 / z2; not w.r.t. pw)."""
    return _WeightedSupCon.apply(z1, z2, pw, enable, float(temperature), bool(in_mode))


class GraphRunner:
    """fwd + bwd of one loss call as ONE CUDA-graph replay, for fixed (n, d) and hyper-parameters.

    The reference's batch sizes (N = 60 .. 512) are launch-bound: 5 kernels + 2 memsets per call.  The graph holds
    the forward and, with an upstream gradient of 1, the backward (the loss is a scalar, so d loss / d z only scales);
    inputs are copied into static buffers.  Labels form only (no tri-state mask)."""

    def __init__(self, n: int, d: int, device, temperature: float, gamma: float, mode: int, correct_grad: bool,
                 use_tc: bool):
        self.key = (n, d, str(device), float(temperature), float(gamma), int(mode), bool(correct_grad), bool(use_tc))
        self.n, self.d = n, d
        g = torch.Generator(device="cpu").manual_seed(0)
        warm = torch.nn.functional.normalize(torch.randn(2, n, d, generator=g), dim=2).to(device)
        self.z1, self.z2 = warm[0].contiguous(), warm[1].contiguous()
        self.labels = torch.arange(n, dtype=torch.int32, device=device)
        self.g = torch.ones(1, dtype=torch.float32, device=device)
        hp = (float(temperature), float(gamma), int(mode))
        fused = not use_tc and int(mode) != nat.MODE_EXCL and fused_fits([(n, d)], device)

        def body_fused():
            N = 2 * n
            z = torch.cat([self.z1, self.z2], dim=0)
            lab2 = torch.cat([self.labels, self.labels])
            ws = torch.empty(N * 8 + 8, dtype=torch.float32, device=device)
            dz = torch.empty(N, d, dtype=torch.float32, device=device)
            probs = (nat.ProblemF32 * 1)()
            _fused_problem(probs[0], z, lab2, N, d, hp[0], hp[1], hp[2], correct_grad, ws, dz)
            if not _fused_launch(probs, 1, ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream), device):
                raise nat.SpclError("the cooperative small-batch launch was declined on this context: set "
                                    "SPCL_FUSED_SMALL=0 to capture the staged route")
            self._keep = (z, lab2, ws)
            return ws[N * 8 + 4:N * 8 + 8], ws[N * 4:N * 8].view(4, N), dz

        def body():
            if fused:
                return body_fused()
            scalars, row_stats, zpack, labels_full, sig = supcon_fwd._init_fn(
                self.z1, self.z2, self.labels, None, hp[0], hp[1], hp[2], bool(correct_grad), bool(use_tc))
            dz = supcon_bwd._init_fn(self.g, zpack, labels_full, sig, None, row_stats, scalars, hp[0], hp[1], hp[2],
                                     bool(use_tc), n, d)
            return scalars, row_stats, dz

        with torch.cuda.device(device):
            side = torch.cuda.Stream(device=device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                for _ in range(2):
                    body()
            torch.cuda.current_stream(device).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = body()

    def run(self, z1: Tensor, z2: Tensor, labels: Tensor):
        """-> static (scalars[4], row_stats[4, n_pad], dz[2n, d] for an upstream gradient of 1); overwritten by the
        next run."""
        self.z1.copy_(z1)
        self.z2.copy_(z2)
        self.labels.copy_(labels)
        with torch.cuda.device(self.z1.device):
            self.graph.replay()
        return self.out


class _GraphedSupCon(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z1, z2, labels, runner):
        scalars, row_stats, dz = runner.run(z1.detach(), z2.detach(), labels)
        ctx.dz = dz.clone()                       # the runner's buffers belong to its next replay
        ctx.n = z1.shape[0]
        ctx.set_materialize_grads(False)
        row_stats = row_stats.clone()
        ctx.mark_non_differentiable(row_stats)
        return scalars.clone(), row_stats

    @staticmethod
    def backward(ctx, g_scalars, _g_stats):
        if g_scalars is None:
            return None, None, None, None
        dz = ctx.dz * g_scalars[0]
        return dz[:ctx.n], dz[ctx.n:], None, None


def supcon_fwd_graphed(z1, z2, labels, runner: GraphRunner):
    return _GraphedSupCon.apply(z1, z2, labels, runner)


# --------------------------------------------------------------------------------------------------
# L2 normalise
# --------------------------------------------------------------------------------------------------
_DTYPES = {torch.float32: nat.DTYPE_F32, torch.bfloat16: nat.DTYPE_BF16, torch.float16: nat.DTYPE_F16}


def _odi(shape, dim: int):
    dim = dim % len(shape)
    return math.prod(shape[:dim]), shape[dim], math.prod(shape[dim + 1:])


@torch.library.custom_op("spcl::l2norm_fwd", mutates_args=(), device_types="cuda")
@on_device_of(0)
def l2norm_fwd(x: Tensor, dim: int, eps: float) -> Tuple[Tensor, Tensor]:
    _require_cuda(x)
    if x.dtype not in _DTYPES:
        raise TypeError(f"unsupported dtype {x.dtype}")
    x = x.contiguous()
    outer, d, inner = _odi(x.shape, dim)
    y = torch.empty_like(x)
    inv = torch.empty(outer * inner, dtype=torch.float32, device=x.device)
    if x.numel():
        nat.call("spcl_l2norm_fwd", _ptr(x), _ptr(y), _ptr(inv), _DTYPES[x.dtype], outer, d, inner, float(eps),
                 _stream(x))
    return y, inv


@l2norm_fwd.register_fake
def _(x, dim, eps):
    outer, d, inner = _odi(x.shape, dim)
    return torch.empty_like(x, memory_format=torch.contiguous_format), x.new_empty(outer * inner,
                                                                                   dtype=torch.float32)


@torch.library.custom_op("spcl::l2norm_bwd", mutates_args=(), device_types="cuda")
@on_device_of(1)
def l2norm_bwd(gy: Tensor, y: Tensor, inv_norm: Tensor, dim: int) -> Tensor:
    _require_cuda(gy, y, inv_norm)
    gy = gy.contiguous().to(y.dtype)
    outer, d, inner = _odi(y.shape, dim)
    gx = torch.empty_like(y)
    if y.numel():
        nat.call("spcl_l2norm_bwd", _ptr(gy), _ptr(y), _ptr(inv_norm), _ptr(gx), _DTYPES[y.dtype], outer, d, inner,
                 _stream(y))
    return gx


@l2norm_bwd.register_fake
def _(gy, y, inv_norm, dim):
    return torch.empty_like(y)


def _l2_setup(ctx, inputs, output):
    _, dim, _ = inputs
    y, inv = output
    ctx.save_for_backward(y, inv)
    ctx.dim = dim
    ctx.set_materialize_grads(False)


def _l2_backward(ctx, gy, g_inv):
    if gy is None:
        return None, None, None
    y, inv = ctx.saved_tensors
    return l2norm_bwd(gy, y, inv, ctx.dim), None, None


l2norm_fwd.register_autograd(_l2_backward, setup_context=_l2_setup)


# --------------------------------------------------------------------------------------------------
# dense-contrast front end: pool + normalise + point gather / rows reshape in one pass
# --------------------------------------------------------------------------------------------------
class _DenseRows(torch.autograd.Function):
    """x [B, C, H, W] fp32 -> unit rows [B*P, C] (heads.py:109-115 + infonce.py:233-241 / comparable.py:398-404)."""

    @staticmethod
    @on_device_of(1)
    def forward(ctx, x, points, ph, pw, eps, pool_max=False):
        _require_cuda(x, points)
        x = x.contiguous()
        B, C, H, W = x.shape
        P = ph * pw if points is None else points.shape[1]
        y = torch.empty(B * P, C, dtype=torch.float32, device=x.device)
        inv = torch.empty(B * P, dtype=torch.float32, device=x.device)
        argmax = torch.empty(B * P, C, dtype=torch.int32, device=x.device) if pool_max else None
        if y.numel():
            if pool_max:
                nat.call("spcl_dense_rows_max_fwd", _ptr(x), _ptr(points), B, C, H, W, ph, pw, P, float(eps), _ptr(y),
                         _ptr(inv), _ptr(argmax), _stream(x))
            else:
                nat.call("spcl_dense_rows_fwd", _ptr(x), _ptr(points), B, C, H, W, ph, pw, P, float(eps), _ptr(y),
                         _ptr(inv), _stream(x))
        ctx.save_for_backward(y, inv, points, argmax)
        ctx.geom = (B, C, H, W, ph, pw, P)
        ctx.normalized = eps >= 0
        ctx.mark_non_differentiable(inv)
        return y, inv

    @staticmethod
    def backward(ctx, gy, _g_inv):
        if gy is None:
            return None, None, None, None, None, None
        y, inv, points, argmax = ctx.saved_tensors
        B, C, H, W, ph, pw, P = ctx.geom
        gx = torch.empty(B, C, H, W, dtype=torch.float32, device=y.device)
        if gx.numel():
            gy = gy.float().contiguous()
            if argmax is not None:
                g_pooled = l2norm_bwd._init_fn(gy, y, inv, 1) if ctx.normalized else gy   # d(loss)/d(pooled value)
                nat.call("spcl_dense_rows_max_bwd", _ptr(g_pooled), _ptr(argmax), B, C, H, W, P, _ptr(gx), _stream(y))
            elif not ctx.normalized:
                nat.call("spcl_dense_rows_bwd", _ptr(gy), _ptr(points), B, C, H, W, ph, pw, P, _ptr(gx), _stream(y))
            else:
                # the normalise backward is folded into the pooling backward's load phase (one launch, no g_pooled)
                nat.call("spcl_dense_rows_bwd_fused", _ptr(gy), _ptr(y), _ptr(inv), _ptr(points), B, C, H, W, ph, pw, P,
                         _ptr(gx), _stream(y))
        return gx, None, None, None, None, None


def dense_rows(x: Tensor, spatial_size, points: Optional[Tensor] = None, eps: float = 1e-12,
               pool: str = "avg", normalize: bool = True) -> Tensor:
    """Unit-norm anchor rows from a dense projector output.

    ``x``: ``[B, C, H, W]`` (any float dtype; computed in fp32).  ``spatial_size``: the adaptive-pool target
    ``(ph, pw)`` (``None`` = no pooling); ``pool``: ``"avg"`` (``AdaptiveAvgPool2d``, the reference's default) or
    ``"max"`` (``AdaptiveMaxPool2d``, ``pool_name="adaptive_max"``).  ``points``: ``None`` for every pooled pixel (rows ordered
    ``(b, i, j)``) or an integer tensor ``[B, P]`` of flat pooled coordinates ``i * pw + j`` (rows ordered
    ``(b, p)``).  ``normalize=False`` returns the pooled rows without the L2 normalisation.  Differentiable with
    respect to ``x``."""
    if x.dim() != 4:
        raise ValueError(f"expected [B, C, H, W], got {tuple(x.shape)}")
    _require_cuda(x)
    H, W = x.shape[2:]
    if spatial_size is None:
        ph, pw = H, W
    elif isinstance(spatial_size, int):
        ph = pw = int(spatial_size)
    else:
        ph, pw = (int(v) for v in spatial_size)
    if ph > H or pw > W or ph <= 0 or pw <= 0:
        raise nat.SpclError(f"spatial_size {(ph, pw)} must pool down from {(H, W)}")
    if points is not None:
        if points.dim() != 2 or points.shape[0] != x.shape[0]:
            raise ValueError(f"points must be [B, P], got {tuple(points.shape)} for B = {x.shape[0]}")
        if not points.is_cuda:                       # host-made coordinates: range-check without a device sync
            if points.numel() and (int(points.min()) < 0 or int(points.max()) >= ph * pw):
                raise IndexError(f"point coordinate outside the {ph} x {pw} pooled grid")
        points = points.to(device=x.device, dtype=torch.int32).contiguous()
    if pool not in ("avg", "max"):
        raise ValueError(f"pool must be 'avg' or 'max', got {pool!r}")
    # normalize=False: the pooled rows themselves (negative eps is the kernels' "do not normalise" switch)
    y, _ = _DenseRows.apply(x.float(), points, ph, pw, float(eps) if normalize else -1.0, pool == "max")
    return y


# --------------------------------------------------------------------------------------------------
# grouped launch: the K meta-label losses of one step share every kernel launch
# --------------------------------------------------------------------------------------------------
class _GroupSupCon(torch.autograd.Function):
    """fp32 label-form problems; inputs z1_0, z2_0, z1_1, z2_1, ...; outputs one scalars[4] per problem."""

    @staticmethod
    @on_device_of(3)
    def forward(ctx, meta, labels, *views):
        K = len(meta)
        dev = views[0].device
        st = _stream(views[0])
        shapes = [tuple(views[2 * k].shape) for k in range(K)]
        # one workspace for every problem's acc [N,4] | row_stats [4,N] | partials [4] | scalars [4]
        offs, total = [], 0
        for n, d in shapes:
            offs.append(total)
            total += 2 * n * 8 + 8
        ws = torch.empty(total, dtype=torch.float32, device=dev)
        probs = (nat.ProblemF32 * K)()
        zs, scal = [], []
        for k, ((n, d), (temperature, gamma, mode, cg)) in enumerate(zip(shapes, meta)):
            N = 2 * n
            z = torch.cat([views[2 * k], views[2 * k + 1]], dim=0)          # contrast_loss3.py:26
            zs.append(z)
            scal.append(ws[offs[k] + N * 8 + 4:offs[k] + N * 8 + 8])
            base = ws.data_ptr() + offs[k] * 4          # acc | row_stats | partials | scalars (pointer arithmetic:
            q = probs[k]                                # every tensor slice costs ~3 us on the host)
            q.z, q.n_total, q.d, q.ldz, q.labels = z.data_ptr(), N, d, z.stride(0), labels[k].data_ptr()
            q.inv_tau, q.gamma, q.mode, q.correct_grad = 1.0 / float(temperature), float(gamma), int(mode), int(bool(cg))
            q.acc, q.row_stats, q.stats_stride = base, base + N * 16, N
            q.partials, q.scalars = base + N * 32, base + N * 32 + 16
        ctx.fused = fused_fits(shapes, dev)
        if ctx.fused:
            # forward AND backward (upstream gradient 1) in one cooperative launch; backward() only scales
            ctx.dz = [torch.empty(2 * n, d, dtype=torch.float32, device=dev) for n, d in shapes]
            for k, dz in enumerate(ctx.dz):
                probs[k].dz, probs[k].lddz = dz.data_ptr(), dz.stride(0)
            ctx.fused = _fused_launch(probs, K, st, dev)
            if not ctx.fused:
                ctx.dz = None
        if not ctx.fused:
            nat.call("spcl_supcon_group_fwd_f32", ctypes.byref(probs), K, st)
        ctx.probs, ctx.keep, ctx.shapes = probs, (ws, zs, labels), shapes
        return tuple(scal)

    @staticmethod
    def backward(ctx, *g_scalars):
        probs, shapes = ctx.probs, ctx.shapes
        K = len(shapes)
        ws, zs, labels = ctx.keep
        if ctx.fused:
            grads = []
            for k, (n, d) in enumerate(shapes):
                if g_scalars[k] is None:
                    grads += [None, None]
                else:
                    dz = ctx.dz[k] * g_scalars[k][0]
                    grads += [dz[:n], dz[n:]]
            return (None, None, *grads)
        st = _stream(ws)
        grads, hold = [], []
        for k, (n, d) in enumerate(shapes):
            g = g_scalars[k]
            g0 = (torch.zeros(1, dtype=torch.float32, device=ws.device) if g is None else g.float().contiguous()[0:1])
            dz = torch.empty(2 * n, d, dtype=torch.float32, device=ws.device)
            hold.append(g0)
            probs[k].grad_out, probs[k].dz, probs[k].lddz = g0.data_ptr(), dz.data_ptr(), d
            grads += [dz[:n], dz[n:]]
        nat.call("spcl_supcon_group_bwd_f32", ctypes.byref(probs), K, st)
        return (None, None, *grads)


def supcon_group_f32(views, labels, meta):
    """``views``: [(z1, z2), ...] fp32 unit rows; ``labels``: per problem int32 [2n] (both views, already tiled);
    ``meta``: per problem (temperature, gamma, mode, correct_grad).  -> list of scalars[4] tensors."""
    K = len(views)
    if not 1 <= K <= nat.MAX_GROUP:
        raise nat.SpclError(f"a group holds 1..{nat.MAX_GROUP} problems, got {K}")
    flat = []
    for (z1, z2), lab in zip(views, labels):
        _require_cuda(z1, z2, lab)
        if z1.shape != z2.shape or z1.dim() != 2:
            raise AssertionError((tuple(z1.shape), tuple(z2.shape)))
        if z1.shape[1] > nat.MAX_D:
            raise nat.SpclError(f"embedding width {z1.shape[1]} > {nat.MAX_D} is not supported by this build")
        if lab.dtype != torch.int32 or lab.shape != (2 * z1.shape[0],):
            raise TypeError("labels must be int32[2n]")
        flat += [z1.float().contiguous(), z2.float().contiguous()]
    return list(_GroupSupCon.apply(tuple(meta), tuple(labels), *flat))


class GroupGraphRunner:
    """Grouped forward + backward (upstream gradients of 1) of K fp32 label-form problems as ONE CUDA-graph replay, for
    fixed shapes and hyper-parameters: 7 kernels per training step's losses instead of K x 13 launches.  Inputs are
    copied into static buffers with two ``_foreach_copy_`` calls; outputs live in static buffers until the next run."""

    def __init__(self, shapes, meta, device):
        self.shapes, self.meta, self.K = list(shapes), list(meta), len(shapes)
        g = torch.Generator(device="cpu").manual_seed(0)
        self.z, self.lab, self.dz = [], [], []
        total = 0
        for n, d in self.shapes:
            warm = torch.nn.functional.normalize(torch.randn(2 * n, d, generator=g), dim=1).to(device)
            self.z.append(warm.contiguous())
            self.lab.append(torch.arange(2 * n, dtype=torch.int32, device=device) % n)
            self.dz.append(torch.empty(2 * n, d, dtype=torch.float32, device=device))
            total += 2 * n * 8 + 8
        self.ws = torch.empty(total, dtype=torch.float32, device=device)
        self.scalars = torch.empty(self.K, 4, dtype=torch.float32, device=device)
        self.ones = torch.ones(1, dtype=torch.float32, device=device)
        self.probs = (nat.ProblemF32 * self.K)()
        off = 0
        for k, ((n, d), (temperature, gamma, mode, cg)) in enumerate(zip(self.shapes, self.meta)):
            N, q, base = 2 * n, self.probs[k], self.ws[off:]
            off += N * 8 + 8
            q.z, q.n_total, q.d, q.ldz, q.labels = self.z[k].data_ptr(), N, d, d, self.lab[k].data_ptr()
            q.inv_tau, q.gamma, q.mode, q.correct_grad = 1.0 / float(temperature), float(gamma), int(mode), int(bool(cg))
            q.acc, q.row_stats, q.stats_stride = base.data_ptr(), base[N * 4:].data_ptr(), N
            q.partials, q.scalars = base[N * 8:].data_ptr(), self.scalars[k].data_ptr()
            q.grad_out, q.dz, q.lddz = self.ones.data_ptr(), self.dz[k].data_ptr(), d
        # destination views of the per-step input copies: z1 | z2 halves, labels twice
        self.z_dst = [h for (n, _), z in zip(self.shapes, self.z) for h in (z[:n], z[n:])]
        self.lab_dst = [h for (n, _), l in zip(self.shapes, self.lab) for h in (l[:n], l[n:])]

        self.fused = fused_fits(self.shapes, device)

        def body():
            st = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
            if self.fused:                       # the whole step's losses and their gradients: ONE kernel
                self.fused = _fused_launch(self.probs, self.K, st, device)
                if self.fused:
                    return
            nat.call("spcl_supcon_group_fwd_f32", ctypes.byref(self.probs), self.K, st)
            nat.call("spcl_supcon_group_bwd_f32", ctypes.byref(self.probs), self.K, st)

        with torch.cuda.device(device):
            side = torch.cuda.Stream(device=device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                for _ in range(2):
                    body()
            torch.cuda.current_stream(device).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                body()

    def run(self, flat_views, labels):
        torch._foreach_copy_(self.z_dst, list(flat_views))
        torch._foreach_copy_(self.lab_dst, [l for l in labels for _ in (0, 1)])
        with torch.cuda.device(self.ws.device):
            self.graph.replay()
        return self.scalars, self.dz


class _GroupGraphed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner, labels, *views):
        scalars, dz = runner.run([v.detach() for v in views], labels)
        ctx.dz = [t.clone() for t in dz]                 # the runner's buffers belong to its next replay
        ctx.ns = [n for n, _ in runner.shapes]
        return scalars.clone()

    @staticmethod
    def backward(ctx, g_scalars):
        grads = []
        for k, (dz, n) in enumerate(zip(ctx.dz, ctx.ns)):
            dz = dz * g_scalars[k, 0]
            grads += [dz[:n], dz[n:]]
        return (None, None, *grads)


def supcon_group_f32_graphed(views, labels, runner: GroupGraphRunner):
    """``labels``: per problem int32 [n] (one view; tiled inside).  -> scalars [K, 4]."""
    flat = []
    for (z1, z2) in views:
        _require_cuda(z1, z2)
        flat += [z1.float(), z2.float()]
    return _GroupGraphed.apply(runner, tuple(labels), *flat)
