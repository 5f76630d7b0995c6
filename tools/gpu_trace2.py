"""Development tool: stats-kernel timeline of CTA 0 for any number of epilogue warps (needs a -DSPCL_TRACE=1 build).

    SPCL_B200_LIB=.../libspcl_trace.so python tools/gpu_trace2.py [n] [d] [first_tile] [ntiles]
Per tile: producer slot-free stamp, MMA ready / issued, and per epilogue warp (S visible, done).
"""
import ctypes
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import spcl_b200  # noqa: E402,F401
from spcl_b200 import ops  # noqa: E402
from spcl_b200._native import lib  # noqa: E402

ROLES = 16


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    t0 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    nt = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    g = torch.Generator().manual_seed(0)
    base = torch.randn(n, d, generator=g)
    z1 = torch.nn.functional.normalize(base + 0.7 * torch.randn(n, d, generator=g), dim=1).cuda()
    z2 = torch.nn.functional.normalize(base + 0.7 * torch.randn(n, d, generator=g), dim=1).cuda()
    lab = torch.arange(n).int().cuda()
    h = lib()
    h.spcl_debug_set_trace.argtypes = [ctypes.c_void_p]
    tr = torch.zeros(ROLES * 64 * 4, dtype=torch.int64, device="cuda")
    for _ in range(2):
        ops.supcon_fwd(z1, z2, lab, None, 0.07, 8.0, 0, False, True)
    torch.cuda.synchronize()
    h.spcl_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
    ops.supcon_fwd(z1, z2, lab, None, 0.07, 8.0, 0, False, True)
    torch.cuda.synchronize()
    h.spcl_debug_set_trace(None)
    t = tr.view(ROLES, 64, 4).cpu().numpy().astype("int64")
    basec = t[t > 0].min()
    r = lambda v: (int(v - basec) if v > 0 else -1)
    nw = max(w for w in range(ROLES - 2) if t[2 + w].any()) + 1
    print(f"epilogue warps seen: {nw}")
    for i in range(t0, min(64, t0 + nt)):
        per = "  ".join(f"w{w}:{r(t[2 + w, i, 0])}/{r(t[2 + w, i, 1])}" for w in range(nw) if t[2 + w, i, 0] > 0)
        print(f"{i:3d} | prod {r(t[0, i, 0]):7d} | mma {r(t[1, i, 0]):7d} {r(t[1, i, 1]):7d} | {per}")


if __name__ == "__main__":
    main()
