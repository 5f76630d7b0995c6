"""Development tool: device time of the forward's parts at cfg3 (prepare | stats pass A | sp pass B + row_finalize) and of
the backward, per labelling and mode.   python tools/gpu_fwd_parts.py"""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import spcl_b200  # noqa: E402,F401
from spcl_b200 import _native as nat  # noqa: E402
from spcl_b200 import ops  # noqa: E402
from spcl_b200.ops import _ptr, _stream  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(500_000)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3


def main():
    n, d = 16384, 128
    N = 2 * n
    g = torch.Generator().manual_seed(0)
    base = torch.randn(n, d, generator=g)
    z1 = torch.nn.functional.normalize(base + 0.7 * torch.randn(n, d, generator=g), dim=1).cuda()
    z2 = torch.nn.functional.normalize(base + 0.7 * torch.randn(n, d, generator=g), dim=1).cuda()
    dev = z1.device
    st = _stream(z1)
    for kind in ("self", "slice"):
        lab = (torch.arange(n) if kind == "self" else torch.arange(n) // 1024).int().cuda()
        for mode, name in ((0, "none"), (1, "hard"), (2, "soft")):
            gamma = 10.0
            scalars, row_stats, zpack, labels_full, sig = ops.supcon_fwd(z1, z2, lab, None, 0.07, gamma, mode, False, True)
            acc = torch.zeros(N, 4, dtype=torch.float32, device=dev)
            partials = torch.zeros(4, dtype=torch.float32, device=dev)
            rs = torch.empty(4, N, dtype=torch.float32, device=dev)
            inv_tau = 1.0 / 0.07

            def stats():
                acc.zero_()
                nat.call("spcl_supcon_stats_part_bf16", _ptr(zpack), N, N, d, _ptr(labels_full), _ptr(sig), 0, 1, inv_tau,
                         mode, _ptr(acc), st)
            stats()
            acc0 = acc.clone()

            def finish():
                acc.copy_(acc0)
                nat.call("spcl_supcon_fwd_finish_bf16", _ptr(zpack), N, N, d, _ptr(labels_full), _ptr(sig), 0, N, inv_tau,
                         gamma, mode, _ptr(acc), _ptr(rs), _ptr(partials), st)
            gone = torch.ones(1, device=dev)
            bwd = lambda: ops.supcon_bwd(gone, zpack, labels_full, sig, None, row_stats, scalars, 0.07, gamma, mode, True, n, d)
            t_copy = timeit(lambda: acc.copy_(acc0))
            print(f"{kind:5s} {name:4s} | stats {timeit(stats):7.1f} us | sp+finalize {timeit(finish) - t_copy:7.1f} us | "
                  f"bwd {timeit(bwd):7.1f} us", flush=True)


if __name__ == "__main__":
    main()
