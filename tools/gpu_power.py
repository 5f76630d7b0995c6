"""Development tool: SM clock / power draw while one op of the bf16 path runs in a tight loop (is the kernel
power-capped?).   python tools/gpu_power.py [fwd|bwd|both] [seconds]"""
import pathlib
import subprocess
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import spcl_b200  # noqa: E402,F401
from spcl_b200 import ops  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "both"
    secs = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
    n, d = 16384, 128
    g = torch.Generator().manual_seed(0)
    base = torch.randn(n, d, generator=g)
    z1 = torch.nn.functional.normalize(base + 0.7 * torch.randn(n, d, generator=g), dim=1).cuda()
    z2 = torch.nn.functional.normalize(base + 0.7 * torch.randn(n, d, generator=g), dim=1).cuda()
    lab = torch.arange(n).int().cuda()
    fwd = lambda: ops.supcon_fwd(z1, z2, lab, None, 0.07, 8.0, 0, False, True)
    scalars, row_stats, zpack, labels_full, sig = fwd()
    gone = torch.ones(1, device="cuda")
    bwd = lambda: ops.supcon_bwd(gone, zpack, labels_full, sig, None, row_stats, scalars, 0.07, 8.0, 0, True, n, d)
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        if which not in (name, "both"):
            continue
        torch.cuda.synchronize()
        time.sleep(1.0)
        smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active",
                                "--format=csv,noheader,nounits", "-lms", "20", "-i", "0"], stdout=subprocess.PIPE, text=True)
        t0 = time.time()
        iters = 0
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        while time.time() - t0 < secs:
            for _ in range(50):
                fn()
            iters += 50
            torch.cuda.synchronize()
        e.record()
        torch.cuda.synchronize()
        smi.terminate()
        rows = [l.split(",") for l in smi.stdout.read().strip().splitlines() if l.count(",") == 2]
        half = rows[len(rows) // 2:]                    # steady state
        clk = sorted(int(r[0]) for r in half)
        pw = sorted(float(r[1]) for r in half)
        reasons = sorted({r[2].strip() for r in half})
        print(f"{name}: {s.elapsed_time(e) / iters * 1e3:7.1f} us/op over {iters} ops | sm clock median {clk[len(clk) // 2]} MHz "
              f"(min {clk[0]}, max {clk[-1]}) | power median {pw[len(pw) // 2]:.0f} W (max {pw[-1]:.0f}) | throttle {reasons}",
              flush=True)


if __name__ == "__main__":
    main()
