"""Development tool: CTA-pair kernels against the single-CTA kernels (debug flag 32 / 64 force the latter).

    python tools/gpu_pair_check.py [time]
Both compute the same arithmetic on the same bf16 operands, so outputs must agree to fp32 summation order.
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
from spcl_b200.workloads import acdc_meta_labels, make_views  # noqa: E402

FORCE_OLD = 32 | 64


def run(n, d, kind, mode, gamma, flags):
    h = lib()
    h.spcl_debug_set_flags.argtypes = [ctypes.c_int]
    if kind == "self":
        labels = torch.arange(n)
        g = torch.Generator().manual_seed(0)
        base = torch.randn(n, d, generator=g)
        z1 = torch.nn.functional.normalize(base + 0.7 * torch.randn(n, d, generator=g), dim=1)
        z2 = torch.nn.functional.normalize(base + 0.7 * torch.randn(n, d, generator=g), dim=1)
    elif kind == "slice":
        labels = torch.arange(n) // max(1, n // 16)
        z1, z2 = make_views(labels, d, sigma=0.7, seed=0)
    else:
        labels = acdc_meta_labels(n)[kind]
        z1, z2 = make_views(labels, d, sigma=0.7, seed=0)
    z1, z2, lab = z1.cuda(), z2.cuda(), labels.int().cuda()
    h.spcl_debug_set_flags(flags)
    scalars, row_stats, zpack, labels_full, sig = ops.supcon_fwd(z1, z2, lab, None, 0.07, gamma, mode, False, True)
    gone = torch.ones(1, device="cuda")
    dz = ops.supcon_bwd(gone, zpack, labels_full, sig, None, row_stats, scalars, 0.07, gamma, mode, True, n, d)
    torch.cuda.synchronize()
    h.spcl_debug_set_flags(0)
    return scalars.clone(), row_stats.clone(), dz.clone()


def timeit(n, d, flags, mode=2, gamma=8.0, iters=20):
    h = lib()
    h.spcl_debug_set_flags.argtypes = [ctypes.c_int]
    g = torch.Generator().manual_seed(0)
    base = torch.randn(n, d, generator=g)
    z1 = torch.nn.functional.normalize(base + 0.7 * torch.randn(n, d, generator=g), dim=1).cuda()
    z2 = torch.nn.functional.normalize(base + 0.7 * torch.randn(n, d, generator=g), dim=1).cuda()
    lab = torch.arange(n).int().cuda()
    h.spcl_debug_set_flags(flags)
    out = ops.supcon_fwd(z1, z2, lab, None, 0.07, gamma, mode, False, True)
    scalars, row_stats, zpack, labels_full, sig = out
    gone = torch.ones(1, device="cuda")
    res = {}
    for name, fn in (("fwd", lambda: ops.supcon_fwd(z1, z2, lab, None, 0.07, gamma, mode, False, True)),
                     ("bwd", lambda: ops.supcon_bwd(gone, zpack, labels_full, sig, None, row_stats, scalars, 0.07,
                                                    gamma, mode, True, n, d))):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        res[name] = s.elapsed_time(e) / iters * 1e3
    h.spcl_debug_set_flags(0)
    return res


def main():
    ok = True
    cases = [(128, 128, "partition", 2, 5.0), (256, 128, "patient", 2, 5.0), (384, 128, "self", 0, 1e6),
             (1024, 128, "slice", 1, 6.0), (2048, 128, "self", 2, 8.0), (640, 100, "cycle", 2, 4.0),
             (4096, 128, "slice", 2, 8.0)]
    for n, d, kind, mode, gamma in cases:
        a = run(n, d, kind, mode, gamma, FORCE_OLD)
        b = run(n, d, kind, mode, gamma, 0)
        ds = (a[0] - b[0]).abs().max().item()
        dr = (a[1] - b[1]).abs().max().item() / max(a[1].abs().max().item(), 1e-30)
        dd = (a[2] - b[2]).abs().max().item() / max(a[2].abs().max().item(), 1e-30)
        good = ds < 1e-5 and dr < 1e-5 and dd < 1e-4 and torch.isfinite(b[2]).all().item()
        ok &= good
        print(f"n={n} d={d} {kind} mode={mode}: scalars diff {ds:.2e} row_stats rel {dr:.2e} dz rel {dd:.2e} "
              f"{'OK' if good else 'MISMATCH'}", flush=True)
    if len(sys.argv) > 1:
        for n in (16384,):
            old = timeit(n, 128, FORCE_OLD)
            new = timeit(n, 128, 0)
            print(f"N={2 * n}: single-CTA fwd {old['fwd']:.1f} us bwd {old['bwd']:.1f} us | pair fwd {new['fwd']:.1f} us "
                  f"bwd {new['bwd']:.1f} us")
    print("ALL OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
