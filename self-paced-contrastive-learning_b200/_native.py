"""ctypes binding of the C ABI declared in ``include/spcl.h`` (``libspcl_b200.so``).

There is no CPU or PyTorch fallback: if the shared library is missing or a call fails,
the error is raised to the caller.
"""
from __future__ import annotations

import ctypes
import os
import pathlib
import subprocess
import threading

_PKG_DIR = pathlib.Path(__file__).resolve().parent
LIB_PATH = pathlib.Path(os.environ.get("SPCL_B200_LIB", _PKG_DIR / "libspcl_b200.so"))   # override: kernel A/B builds
BUILD_SCRIPT = _PKG_DIR / "csrc" / "build.sh"

MODE_NONE, MODE_HARD, MODE_SOFT, MODE_EXCL = 0, 1, 2, 3
ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DRIVER = -1, -2, -3, -4
DTYPE_F32, DTYPE_BF16, DTYPE_F16 = 0, 1, 2
TILE = 128
MAX_D = 256

_c = ctypes
_i64, _i32, _f32, _ptr = _c.c_int64, _c.c_int32, _c.c_float, _c.c_void_p

# name -> argtypes; every entry point of include/spcl.h that returns an error code
SIGNATURES = {
    "spcl_l2norm_fwd": [_ptr, _ptr, _ptr, _c.c_int, _i64, _i64, _i64, _f32, _ptr],
    "spcl_l2norm_bwd": [_ptr, _ptr, _ptr, _ptr, _c.c_int, _i64, _i64, _i64, _ptr],
    "spcl_pack_views_bf16": [_ptr, _ptr, _i64, _i64, _i64, _i64, _ptr, _i64, _ptr],
    "spcl_label_block_sig": [_ptr, _i64, _i64, _ptr, _ptr],
    "spcl_supcon_prepare_bf16": [_ptr, _ptr, _i64, _i64, _i64, _i64, _ptr, _ptr, _i64, _i64, _ptr, _ptr, _ptr, _ptr],
    "spcl_supcon_prepare_raw_bf16": [_ptr, _ptr, _i64, _i64, _i64, _f32, _ptr, _ptr, _i64, _i64, _ptr, _ptr, _ptr, _ptr,
                                     _ptr],
    "spcl_supcon_raw_bwd": [_ptr, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _ptr],
    "spcl_dense_rows_fwd": [_ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _f32, _ptr, _ptr, _ptr],
    "spcl_dense_rows_bwd": [_ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr],
    "spcl_dense_rows_bwd_fused": [_ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr],
    "spcl_dense_rows_max_fwd": [_ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _f32, _ptr, _ptr, _ptr, _ptr],
    "spcl_dense_rows_max_bwd": [_ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr],
    "spcl_supcon_fwd_bf16": [_ptr, _i64, _i64, _i32, _ptr, _ptr, _i64, _i64, _f32, _f32, _c.c_int, _ptr, _ptr,
                             _ptr, _ptr],
    "spcl_supcon_stats_part_bf16": [_ptr, _i64, _i64, _i32, _ptr, _ptr, _i32, _i32, _f32, _c.c_int, _ptr, _ptr],
    "spcl_supcon_fwd_finish_bf16": [_ptr, _i64, _i64, _i32, _ptr, _ptr, _i64, _i64, _f32, _f32, _c.c_int, _ptr, _ptr,
                                    _ptr, _ptr],
    "spcl_supcon_bwd_bf16": [_ptr, _i64, _i64, _i32, _i32, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _f32, _f32,
                             _c.c_int, _ptr, _i64, _ptr, _ptr],
    "spcl_supcon_fwd_f32": [_ptr, _i64, _i32, _i64, _ptr, _ptr, _i64, _i64, _i64, _f32, _f32, _c.c_int, _ptr,
                            _i64, _ptr, _ptr],
    "spcl_supcon_bwd_f32": [_ptr, _i64, _i32, _i64, _ptr, _ptr, _i64, _ptr, _i64, _ptr, _ptr, _i64, _i64, _f32,
                            _f32, _c.c_int, _ptr, _i64, _ptr],
    "spcl_supcon_fwd_f32_split": [_ptr, _i64, _i32, _i64, _ptr, _i64, _i64, _f32, _f32, _c.c_int, _ptr, _ptr, _i64,
                                  _ptr, _ptr],
    "spcl_supcon_bwd_f32_split": [_ptr, _i64, _i32, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _i64, _f32, _f32,
                                  _c.c_int, _ptr, _i64, _ptr],
    "spcl_supcon_fwd_w_f32": [_ptr, _i64, _i32, _i64, _ptr, _i64, _ptr, _c.c_int, _f32, _ptr, _i64, _ptr, _ptr],
    "spcl_supcon_bwd_w_f32": [_ptr, _i64, _i32, _i64, _ptr, _i64, _ptr, _c.c_int, _f32, _ptr, _i64, _ptr, _ptr, _ptr,
                              _i64, _ptr],
    "spcl_supcon_finalize": [_ptr, _i64, _c.c_int, _ptr, _ptr],
    "spcl_supcon_group_fwd_f32": [_ptr, _c.c_int, _ptr],
    "spcl_supcon_group_bwd_f32": [_ptr, _c.c_int, _ptr],
    "spcl_supcon_group_fused_f32": [_ptr, _c.c_int, _ptr],
}
MAX_GROUP = 8


class ProblemF32(_c.Structure):
    """``spcl_problem_f32`` of include/spcl.h (field order and types must match)."""
    _fields_ = [("z", _ptr), ("n_total", _i64), ("d", _i32), ("ldz", _i64), ("labels", _ptr),
                ("inv_tau", _f32), ("gamma", _f32), ("mode", _c.c_int), ("correct_grad", _c.c_int),
                ("acc", _ptr), ("row_stats", _ptr), ("stats_stride", _i64), ("partials", _ptr), ("scalars", _ptr),
                ("grad_out", _ptr), ("dz", _ptr), ("lddz", _i64)]
OTHER_SYMBOLS = ("spcl_version", "spcl_error_string", "spcl_last_cuda_error", "spcl_workspace_bytes",
                 "spcl_supcon_fused_capacity")
WS_ZB, WS_LABELS, WS_SIG, WS_ACC, WS_ROW_STATS, WS_PARTIALS, WS_SCALARS, WS_BWD_ZT = range(8)
ALL_SYMBOLS = tuple(SIGNATURES) + OTHER_SYMBOLS

_lib = None
_lock = threading.Lock()


class SpclError(RuntimeError):
    pass


def build(verbose: bool = False) -> pathlib.Path:
    """Compile the CUDA sources for sm_100a into ``libspcl_b200.so`` (in-tree)."""
    env = dict(os.environ)
    res = subprocess.run(["bash", str(BUILD_SCRIPT)], capture_output=True, text=True, env=env)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise SpclError(f"building {LIB_PATH.name} failed (exit {res.returncode})")
    return LIB_PATH


def lib() -> ctypes.CDLL:
    """The loaded shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not LIB_PATH.exists():
                    raise SpclError(
                        f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        f"(or bash {BUILD_SCRIPT}). spcl_b200 has no CPU/PyTorch fallback.")
                handle = ctypes.CDLL(str(LIB_PATH))
                for name, argtypes in SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.argtypes = argtypes
                    fn.restype = _c.c_int
                handle.spcl_version.restype = _c.c_int
                handle.spcl_workspace_bytes.argtypes = [_c.c_int, _i64, _i32]
                handle.spcl_workspace_bytes.restype = _i64
                handle.spcl_error_string.argtypes = [_c.c_int]
                handle.spcl_error_string.restype = _c.c_char_p
                handle.spcl_last_cuda_error.restype = _c.c_char_p
                handle.spcl_supcon_fused_capacity.restype = _c.c_int
                _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        h = lib()
        msg = h.spcl_error_string(rc).decode()
        if rc == -3:
            msg += ": " + h.spcl_last_cuda_error().decode()
        raise SpclError(f"{what} failed with code {rc}: {msg}")


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args), name)


def workspace_bytes(which: int, n_pad: int, d_pad: int) -> int:
    """``spcl_workspace_bytes``: size of a caller-owned buffer of the tensor-core entry points."""
    n = int(lib().spcl_workspace_bytes(which, n_pad, d_pad))
    if n < 0:
        check(n, "spcl_workspace_bytes")
    return n
