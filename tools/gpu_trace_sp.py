"""Development tool: timeline of sp_kernel's CTA 0 at cfg3 (needs a -DSPCL_TRACE=1 build in SPCL_B200_LIB).

    SPCL_B200_LIB=variants/trace.so python tools/gpu_trace_sp.py [self|slice]
"""
import ctypes
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import spcl_b200  # noqa: E402,F401
from spcl_b200 import _native as nat, ops  # noqa: E402
from spcl_b200.ops import _ptr, _stream  # noqa: E402

n, d = 16384, 128
N = 2 * n
kind = sys.argv[1] if len(sys.argv) > 1 else "self"
g = torch.Generator().manual_seed(0)
base = torch.randn(n, d, generator=g)
z1 = torch.nn.functional.normalize(base + 0.7 * torch.randn(n, d, generator=g), dim=1).cuda()
z2 = torch.nn.functional.normalize(base + 0.7 * torch.randn(n, d, generator=g), dim=1).cuda()
lab = (torch.arange(n) if kind == "self" else torch.arange(n) // 1024).int().cuda()
mode, gamma, inv_tau = 2, 10.0, 1.0 / 0.07
scalars, row_stats, zpack, labels_full, sig = ops.supcon_fwd(z1, z2, lab, None, 0.07, gamma, mode, False, True)
dev, st = z1.device, _stream(z1)
acc = torch.zeros(N, 4, dtype=torch.float32, device=dev)
nat.call("spcl_supcon_stats_part_bf16", _ptr(zpack), N, N, d, _ptr(labels_full), _ptr(sig), 0, 1, inv_tau, mode, _ptr(acc), st)
acc0 = acc.clone()
partials = torch.zeros(4, dtype=torch.float32, device=dev)
rs = torch.empty(4, N, dtype=torch.float32, device=dev)


def finish():
    acc.copy_(acc0)
    nat.call("spcl_supcon_fwd_finish_bf16", _ptr(zpack), N, N, d, _ptr(labels_full), _ptr(sig), 0, N, inv_tau, gamma, mode,
             _ptr(acc), _ptr(rs), _ptr(partials), st)


for _ in range(3):
    finish()
torch.cuda.synchronize()
h = nat.lib()
h.spcl_debug_set_trace.argtypes = [ctypes.c_void_p]
tr = torch.zeros(10 * 64 * 4, dtype=torch.int64, device="cuda")
h.spcl_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
finish()
torch.cuda.synchronize()
h.spcl_debug_set_trace(None)
t = tr.view(10, 64, 4).cpu().numpy().astype("int64")
t0 = t[0, 62, 0]
r = lambda v: (int(v - t0) if v > 0 else -1)
print(f"sp_kernel CTA 0 ({kind} labels): cycles from kernel entry; prologue done {r(t[0, 62, 1])}, exit {r(t[0, 62, 2])}")
print("tile | producer slot free | mma ready / issued | epilogue warps: S visible / done")
for i in range(12):
    if t[0, i, 0] == 0 and t[1, i, 0] == 0:
        break
    wg = i & 1
    per = "  ".join(f"w{wg * 4 + q}: {r(t[2 + wg * 4 + q, i, 0])}/{r(t[2 + wg * 4 + q, i, 1])}" for q in range(4))
    print(f"{i:4d} | {r(t[0, i, 0]):8d} | {r(t[1, i, 0]):8d} {r(t[1, i, 1]):8d} | {per}")
