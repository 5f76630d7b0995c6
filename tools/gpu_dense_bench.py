"""HBM roofline of the dense front-end kernels (device-timed, L2 flushed between runs).
    python tools/gpu_dense_bench.py > gpurun_out/dense_bench.json"""
import json
import pathlib
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import spcl_b200                                            # noqa: E402

PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6534.8


def timed(fn, reps=10, flush_l2=True, primed=False):
    """Median device time of fn() between two events.  ``primed``: a ~0.5 ms spin kernel is queued first, so the host has
    enqueued fn()'s launches before the GPU reaches the start event -- the time the op takes inside a training step,
    where the launch queue is never empty; without it a lone op also shows the host's Python / autograd latency."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    ms = []
    for _ in range(reps):
        if flush_l2:
            flush.zero_()
        if primed:
            torch.cuda._sleep(1_000_000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    return ms[len(ms) // 2]


CASES = [(32, 128, 224, 224, 32, 32, None), (32, 128, 64, 64, 32, 32, None), (32, 128, 32, 32, 32, 32, None),
         (128, 128, 56, 56, 10, 10, 5)]


def measure(cases=CASES, reps=10):
    out = []
    for (B, C, H, W, ph, pw, P) in cases:
        x = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
        pts = None if P is None else spcl_b200.point_coordinates(B, ph, pw, P, seed=0).cuda()
        rows = spcl_b200.ops.dense_rows(x, (ph, pw), pts)
        gy = torch.randn_like(rows)
        # inputs several times the 126 MB L2 need no flush (and a flush leaves 126 MB of dirty lines to write back
        # inside the timed kernel); smaller ones are flushed
        big = x.numel() * 4 > (512 << 20)
        f_ms = timed(lambda: spcl_b200.ops.dense_rows(x.detach(), (ph, pw), pts), reps, flush_l2=not big)
        b_ms = timed(lambda: torch.autograd.grad(rows, x, gy, retain_graph=True), reps, flush_l2=not big)
        fq_ms = timed(lambda: spcl_b200.ops.dense_rows(x.detach(), (ph, pw), pts), reps, flush_l2=not big, primed=True)
        bq_ms = timed(lambda: torch.autograd.grad(rows, x, gy, retain_graph=True), reps, flush_l2=not big, primed=True)
        n_rows = rows.shape[0]
        fwd_bytes = 4 * (B * C * H * W if P is None else n_rows * C * (H // ph + 1) * (W // pw + 1)) + 4 * n_rows * C
        bwd_bytes = 4 * B * C * H * W + 3 * 4 * n_rows * C
        out.append(dict(shape=[B, C, H, W], pooled=[ph, pw], points=P,
                        l2="input > 4 x L2, no flush" if big else "256 MB flush", fwd_ms=f_ms, bwd_ms=b_ms,
                        fwd_gbs=fwd_bytes / f_ms / 1e6, bwd_gbs=bwd_bytes / b_ms / 1e6,
                        fwd_frac=fwd_bytes / f_ms / 1e6 / PEAK, bwd_frac=bwd_bytes / b_ms / 1e6 / PEAK,
                        fwd_queued_ms=fq_ms, bwd_queued_ms=bq_ms, fwd_queued_frac=fwd_bytes / fq_ms / 1e6 / PEAK,
                        bwd_queued_frac=bwd_bytes / bq_ms / 1e6 / PEAK,
                        queued="launch queue primed by a 0.5 ms spin kernel: the op's own device time, as inside a training step"))
        del x, rows, gy
    return dict(peak_gbs=PEAK, cases=out)


if __name__ == "__main__":
    print(json.dumps(measure()))
