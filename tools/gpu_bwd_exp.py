"""Development tool: backward op time at cfg3 under the debug switches of supcon_tc.cu (results are wrong under them).

    SPCL_B200_LIB=variants/x.so python tools/gpu_bwd_exp.py [flags ...]
flags: 1 = bwd epilogue skips the math (ld + st + barriers only), 8 = S issuer also waits for the T.Z commit.
"""
import ctypes
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import spcl_b200  # noqa: E402,F401
from spcl_b200 import ops  # noqa: E402
from spcl_b200._native import lib  # noqa: E402


def main():
    flags = [int(x) for x in sys.argv[1:]] or [0, 1]
    n, d = 16384, 128
    g = torch.Generator().manual_seed(0)
    base = torch.randn(n, d, generator=g)
    z1 = torch.nn.functional.normalize(base + 0.7 * torch.randn(n, d, generator=g), dim=1).cuda()
    z2 = torch.nn.functional.normalize(base + 0.7 * torch.randn(n, d, generator=g), dim=1).cuda()
    lab = torch.arange(n).int().cuda()
    h = lib()
    h.spcl_debug_set_flags.argtypes = [ctypes.c_int]
    scalars, row_stats, zpack, labels_full, sig = ops.supcon_fwd(z1, z2, lab, None, 0.07, 8.0, 2, False, True)
    gone = torch.ones(1, device="cuda")
    name = pathlib.Path(os.environ.get("SPCL_B200_LIB", "default")).name
    for f in flags:
        h.spcl_debug_set_flags(f)
        bwd = lambda: ops.supcon_bwd(gone, zpack, labels_full, sig, None, row_stats, scalars, 0.07, 8.0, 2, True, n, d)
        for _ in range(3):
            bwd()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            bwd()
        e.record()
        torch.cuda.synchronize()
        print(f"{name:12s} flags={f:3d} bwd {s.elapsed_time(e) / 20 * 1e3:7.1f} us", flush=True)
    h.spcl_debug_set_flags(0)


if __name__ == "__main__":
    main()
