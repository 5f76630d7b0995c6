"""Development tool: fwd_kernel<0> timeline summaries under the debug switches of supcon_tc.cu.

    python tools/gpu_exp.py [n] [d] [flags ...]
flags: 1 = epilogue skips the math, 2 = epilogue skips tcgen05.ld, 4 = producer skips the TMA loads,
       16 = 128 x 128 tiles in the stats kernel (default 128 x 256 when d <= 128).
"""
import ctypes
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import spcl_b200  # noqa: E402,F401
from spcl_b200 import ops  # noqa: E402
from spcl_b200._native import lib  # noqa: E402


def summary(name, tr, lo=8, hi=56, both=True):
    t = tr.view(10, 64, 4).cpu().numpy().astype("int64")
    idx = np.arange(lo, hi)
    wg = 2 + (0 if both else 4 * (idx & 1))
    prod = t[0, idx, 0]
    ready, issued = t[1, idx, 0], t[1, idx, 1]
    vis, done = t[wg, idx, 0], t[wg, idx, 1]
    visb, doneb = t[6, idx, 0], t[6, idx, 1]
    cad = (vis[-1] - vis[0]) / (len(idx) - 1)
    print(f"{name}: cadence {cad:7.0f} cyc/tile | prod->ready {np.mean(ready - prod):7.0f} | ready->issued "
          f"{np.mean(issued - ready):6.0f} | issued->visible {np.mean(vis - issued):6.0f} | epi_dur {np.mean(done - vis):6.0f}"
          + (f" | wg1 epi_dur {np.mean(doneb - visb):6.0f}" if both else ""))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    flags = [int(x) for x in sys.argv[3:]] or [0, 1, 2, 3, 32, 33, 34, 35]
    g = torch.Generator().manual_seed(0)
    base = torch.randn(n, d, generator=g)
    z1 = torch.nn.functional.normalize(base + 0.7 * torch.randn(n, d, generator=g), dim=1).cuda()
    z2 = torch.nn.functional.normalize(base + 0.7 * torch.randn(n, d, generator=g), dim=1).cuda()
    lab = torch.arange(n).int().cuda()
    h = lib()
    h.spcl_debug_set_trace.argtypes = [ctypes.c_void_p]
    h.spcl_debug_set_flags.argtypes = [ctypes.c_int]
    tr = torch.zeros(10 * 64 * 4, dtype=torch.int64, device="cuda")
    for f in flags:
        h.spcl_debug_set_flags(f)
        for _ in range(2):
            ops.supcon_fwd(z1, z2, lab, None, 0.07, 8.0, 0, False, True)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            ops.supcon_fwd(z1, z2, lab, None, 0.07, 8.0, 0, False, True)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 5
        tr.zero_()
        h.spcl_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
        ops.supcon_fwd(z1, z2, lab, None, 0.07, 8.0, 0, False, True)
        torch.cuda.synchronize()
        h.spcl_debug_set_trace(None)
        if tr.any().item():
            summary(f"flags={f} fwd op {ms * 1e3:7.1f} us", tr, both=not (f & 16))
        else:
            print(f"flags={f} fwd op {ms * 1e3:7.1f} us (library built without -DSPCL_TRACE=1: no timeline)")
    h.spcl_debug_set_flags(0)


if __name__ == "__main__":
    main()
