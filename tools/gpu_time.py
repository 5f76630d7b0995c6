"""Development tool: fwd / bwd op times of the bf16 path at cfg3 for the library in SPCL_B200_LIB (A/B builds).

    SPCL_B200_LIB=variants/x.so python tools/gpu_time.py [n] [d] [labels: self|slice]
"""
import os
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import spcl_b200  # noqa: E402,F401
from spcl_b200 import ops  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    kind = sys.argv[3] if len(sys.argv) > 3 else "self"
    g = torch.Generator().manual_seed(0)
    base = torch.randn(n, d, generator=g)
    z1 = torch.nn.functional.normalize(base + 0.7 * torch.randn(n, d, generator=g), dim=1).cuda()
    z2 = torch.nn.functional.normalize(base + 0.7 * torch.randn(n, d, generator=g), dim=1).cuda()
    lab = (torch.arange(n) if kind == "self" else torch.arange(n) // 1024).int().cuda()
    out = {}
    for mode, name in ((0, "none"), (2, "soft")):
        fwd = lambda: ops.supcon_fwd(z1, z2, lab, None, 0.07, 8.0, mode, False, True)
        res = fwd()
        scalars, row_stats, zpack, labels_full, sig = res
        gone = torch.ones(1, device="cuda")
        bwd = lambda: ops.supcon_bwd(gone, zpack, labels_full, sig, None, row_stats, scalars, 0.07, 8.0, mode, True, n, d)
        out[name] = (timeit(fwd), timeit(bwd))
    lib = os.environ.get("SPCL_B200_LIB", "default")
    print(f"{pathlib.Path(lib).name:12s} n={n} d={d} labels={kind} | " +
          " | ".join(f"{k}: fwd {v[0]:7.1f} us bwd {v[1]:7.1f} us" for k, v in out.items()), flush=True)


if __name__ == "__main__":
    main()
