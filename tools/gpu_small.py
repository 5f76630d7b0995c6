"""Development tool: latency of the reference's own batch sizes (cfg1 N=128 d=128, cfg2 N=512 d=256, K=3 meta-label
problems) through the public modules: eager launches vs one CUDA-graph replay of all K fwd+bwd.

    python tools/gpu_small.py
"""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import spcl_b200  # noqa: E402
from spcl_b200.workloads import acdc_meta_labels, make_views  # noqa: E402


def bench(fn, iters=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3


def raw_group_times(n, d):
    """GPU time of the K = 3 losses + gradients of a step through the C ABI alone (no Python between the launches that
    matter: 50 queued repetitions between two events): the one cooperative launch vs the 7 launches of the staged route."""
    import ctypes
    from spcl_b200 import _native as nat, ops
    meta = [(0.07, g, nat.MODE_SOFT, True) for g in (5.0, 3.5, 2.0)]
    runner = ops.GroupGraphRunner([(n, d)] * 3, meta, torch.device("cuda"))
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    fused = lambda: nat.call("spcl_supcon_group_fused_f32", ctypes.byref(runner.probs), 3, st)

    def staged():
        nat.call("spcl_supcon_group_fwd_f32", ctypes.byref(runner.probs), 3, st)
        nat.call("spcl_supcon_group_bwd_f32", ctypes.byref(runner.probs), 3, st)
    out = {"staged_us": bench(staged, 50)}
    if runner.fused:
        out["fused_us"] = bench(fused, 50)
        out["graph_replay_us"] = bench(runner.graph.replay, 50)
    return out


def main():
    for name, n, d in (("cfg1", 64, 128), ("cfg2", 256, 256)):
        print(f"{name}: C ABI only, K=3: {raw_group_times(n, d)}", flush=True)
        meta = acdc_meta_labels(n)
        kinds = ("partition", "patient", "cycle")
        gammas = (5.0, 3.5, 2.0)
        probs = []
        for k, g in zip(kinds, gammas):
            z1, z2 = make_views(meta[k], d, sigma=0.7, seed=1)
            crit = spcl_b200.SelfPacedSupConLoss(weight_update="soft", correct_grad=True, check_nan=False, validate=False)
            crit.set_gamma(g)
            probs.append((crit, z1.cuda().requires_grad_(True), z2.cuda().requires_grad_(True), meta[k].int().cuda()))

        def step():
            total = None
            for crit, a, b, lab in probs:
                a.grad = b.grad = None
                l = crit(a, b, target=lab)
                total = l if total is None else total + l
            total.backward()
            return total

        eager = bench(step)
        # one graph for the K problems (static inputs, like torch.cuda.make_graphed_callables does)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(s)
        for _, a, b, _ in probs:
            a.grad = b.grad = None
        try:
            with torch.cuda.graph(g):
                out = step()
            graphed = bench(g.replay)
            gtxt = f"{graphed:7.1f} us (loss {out.item():.5f})"
        except Exception as ex:  # noqa: BLE001
            gtxt = f"capture failed: {type(ex).__name__}: {str(ex)[:120]}"
        # the modules' own cuda_graph=True mode (one replay per loss call, inputs copied into static buffers)
        gprobs = []
        for (crit, a, b, lab), gm in zip(probs, gammas):
            gc = spcl_b200.SelfPacedSupConLoss(weight_update="soft", correct_grad=True, check_nan=False, validate=False,
                                               cuda_graph=True)
            gc.set_gamma(gm)
            gprobs.append((gc, a, b, lab))

        def gstep():
            total = None
            for crit, a, b, lab in gprobs:
                a.grad = b.grad = None
                l = crit(a, b, target=lab)
                total = l if total is None else total + l
            total.backward()
            return total

        modg = bench(gstep)
        print(f"{name}: N={2 * n} d={d} K=3 problems fwd+bwd | eager {eager:7.1f} us | cuda_graph=True modules {modg:7.1f} us"
              f" | whole-step graph replay {gtxt}", flush=True)


if __name__ == "__main__":
    main()
