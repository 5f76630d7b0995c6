"""Development tool: per-tile timeline of CTA 0 (clock64 stamps written by the TRACE hooks in supcon_tc.cu).

    python tools/gpu_trace.py [n] [d]
Roles: 0 producer (slot free), 1 MMA (0 = S operands+buffer ready, 1 = S issued, 2 = T ready, 3 = T.Z issued),
2/3 epilogue warpgroup A/B (0 = S visible, 1 = tile done).
"""
import ctypes
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import spcl_b200  # noqa: E402
from spcl_b200 import ops  # noqa: E402
from spcl_b200._native import lib  # noqa: E402
from spcl_b200.workloads import make_views  # noqa: E402


def dump(name, tr, ntiles=24, both=False):
    t = tr.view(10, 64, 4).cpu().numpy().astype("int64")
    base = t[t > 0].min()
    r = lambda v: (int(v - base) if v > 0 else -1)
    print(f"--- {name}: cycles relative to first stamp (CTA 0)")
    print("tile | prod.free | mma.ready mma.S_issued mma.T_ready mma.TZ_issued | per epilogue warp of the owning warpgroup: "
          "loop-top/S-visible/first-ld/done")
    for i in range(ntiles):
        wgs = (0, 1) if both else (i & 1,)
        per = []
        for wg in wgs:
            for q in range(4):
                w = 2 + wg * 4 + q
                per.append(f"w{wg * 4 + q}: {r(t[w, i, 3])}/{r(t[w, i, 0])}/{r(t[w, i, 2])}/{r(t[w, i, 1])}")
        print(f"{i:4d} | {r(t[0, i, 0]):9d} | {r(t[1, i, 0]):9d} {r(t[1, i, 1]):9d} {r(t[1, i, 2]):9d} {r(t[1, i, 3]):9d} | "
              + "  ".join(per))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    bwd_flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0      # debug switches for the backward pass only
    labels = torch.arange(n)
    g = torch.Generator().manual_seed(0)
    base = torch.randn(n, d, generator=g)
    z1 = torch.nn.functional.normalize(base + 0.7 * torch.randn(n, d, generator=g), dim=1).cuda()
    z2 = torch.nn.functional.normalize(base + 0.7 * torch.randn(n, d, generator=g), dim=1).cuda()
    lab = labels.int().cuda()
    h = lib()
    h.spcl_debug_set_trace.argtypes = [ctypes.c_void_p]
    tr = torch.zeros(10 * 64 * 4, dtype=torch.int64, device="cuda")
    for _ in range(2):
        out = ops.supcon_fwd(z1, z2, lab, None, 0.07, 8.0, 0, False, True)
    torch.cuda.synchronize()
    h.spcl_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
    out = ops.supcon_fwd(z1, z2, lab, None, 0.07, 8.0, 0, False, True)      # mode NONE: only fwd_kernel<0> runs
    torch.cuda.synchronize()
    h.spcl_debug_set_trace(None)
    dump("stats_kernel<256> (both warpgroups work on every tile; wg A shown; T_ready = operands landed, TZ_issued = "
         "S buffer free)", tr, both=(d <= 128))
    scalars, row_stats, zpack, labels_full, sig = out
    gone = torch.ones(1, device="cuda")
    for _ in range(2):
        ops.supcon_bwd(gone, zpack, labels_full, sig, None, row_stats, scalars, 0.07, 8.0, 0, True, n, d)
    torch.cuda.synchronize()
    tr.zero_()
    h.spcl_debug_set_flags.argtypes = [ctypes.c_int]
    h.spcl_debug_set_flags(bwd_flags)
    h.spcl_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
    ops.supcon_bwd(gone, zpack, labels_full, sig, None, row_stats, scalars, 0.07, 8.0, 0, True, n, d)
    torch.cuda.synchronize()
    h.spcl_debug_set_trace(None)
    h.spcl_debug_set_flags(0)
    dump(f"bwd_kernel flags={bwd_flags}", tr, ntiles=40, both=len(sys.argv) > 4)


if __name__ == "__main__":
    main()
