#!/usr/bin/env python
"""Headline benchmark: self-paced SupCon fwd+bwd pairs/sec (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one forward + backward of the self-paced loss over one synthetic batch.
  N = 1 : cfg3  dense pixel contrast, 2 x 16384 anchors, d = 128  (the config the metric is quoted on)
  N > 1 : cfg4  N = 262144 anchors, d = 128, anchor rows sharded over the ranks (torchrun, NCCL)
Prints ONE JSON line (rank 0).  `value` times the device-resident step with CUDA events; `e2e` times the
public module API from pinned host buffers (H2D + step + D2H of the loss); `roofline` is the backward
kernel alone against the measured bf16 peak; `cpu_baseline` is the oracle's dense fp32 port of the
reference on the host cores.  `--impl reference` times that CPU port as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib
import statistics
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

TAU, GAMMA, MODE_NAME = 0.07, 8.0, "soft"
CPU_SAMPLE_N = 2048          # anchors per view of the bounded CPU sample (N = 4096)


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return float(j["bf16_tflops"]), float(j.get("bf16_tflops_sustained", j["bf16_tflops"])), "measured"
    return 1590.0, 1400.0, "fallback"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons while the timed region runs (NVML; same fields as the nvidia-smi recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def cpu_port_time(workload: str, steps: int, warmup: int):
    """Times the oracle's dense fp32 port (same op sequence as the reference) on the host cores."""
    import torch
    from oracle.dense_port import dense_supcon          # the CPU baseline leg may execute oracle/
    from spcl_b200.workloads import make_workload
    torch.set_num_threads(os.cpu_count() or 1)
    z1, z2, labels = make_workload(workload)
    n = CPU_SAMPLE_N
    z1, z2, labels = z1[:n].clone(), z2[:n].clone(), labels[:n].tolist()
    times = []
    for i in range(warmup + steps):
        a = z1.clone().requires_grad_(True)
        b = z2.clone().requires_grad_(True)
        t0 = time.perf_counter()
        out = dense_supcon(a, b, target=labels, temperature=TAU, gamma=GAMMA, mode=MODE_NAME)
        out.loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    N = 2 * n
    return dict(ms=1e3 * statistics.mean(times), best_ms=1e3 * min(times), N=N, cores=torch.get_num_threads(),
                pairs_per_s=N * N / statistics.mean(times),
                sample=f"first {n} anchors/view of {workload} (N={N}, d={z1.shape[1]}), fp32, fwd+autograd bwd")


def run_reference(args, workload, world, rank):
    if rank != 0:
        return
    r = cpu_port_time(workload, args.steps, max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": "supcon_fwd_bwd_pairs_per_sec", "value": r["pairs_per_s"], "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "hyper": {"tau": TAU, "gamma": GAMMA, "mode": MODE_NAME},
                   "note": "reference is CPU PyTorch; timed on a bounded sample of the workload"},
        "cpu_baseline": {"value": r["pairs_per_s"], "unit": "pairs/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["pairs_per_s"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _ncu_traffic(which):
    """dram bytes of the dominant kernel from the newest committed `ncu --set full` summary (profiles/*_ncu_<which>.txt,
    same bench command, cfg3); a profiler figure cannot be taken inside this run.  -> (bytes | None, source)."""
    import pathlib
    import re
    files = sorted((pathlib.Path(__file__).resolve().parent / "profiles").glob(f"r*_ncu_{which}.txt"))
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for f in reversed(files):
        tot, seen = 0.0, 0
        for line in f.read_text().splitlines():
            m = re.match(r"dram__bytes_(read|write)\.sum\s+([0-9.,]+)\s+(\w+)", line)
            if m and m.group(3) in unit:
                tot += float(m.group(2).replace(",", "")) * unit[m.group(3)]
                seen += 1
        if seen >= 2:
            return tot, f"profiles/{f.name}"
    return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = "cfg3_dense_2x16384_d128_simclr" if world == 1 else "cfg4_dense_2x131072_d128_simclr"

    if args.impl == "reference":
        run_reference(args, workload, world, rank)
        return

    import torch
    import torch.distributed as dist

    import spcl_b200
    from spcl_b200 import _native as nat
    from spcl_b200 import ops
    from spcl_b200.workloads import WORKLOADS, make_workload

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        from spcl_b200.distributed import sharded_supcon_loss

    spec = WORKLOADS[workload]
    n_total, d = spec["n"], spec["d"]
    N = 2 * n_total
    n_loc = n_total // world
    # synthetic anchors: every rank generates only its own samples (seed = base + rank)
    if world == 1:
        z1h, z2h, labels_h = make_workload(workload, seed=0)
    else:
        g = torch.Generator().manual_seed(1000 + rank)
        base = torch.randn(n_loc, d, generator=g)
        z1h = torch.nn.functional.normalize(base + 0.7 * torch.randn(n_loc, d, generator=g), dim=1)
        z2h = torch.nn.functional.normalize(base + 0.7 * torch.randn(n_loc, d, generator=g), dim=1)
        labels_h = torch.arange(rank * n_loc, (rank + 1) * n_loc)
    z1h, z2h = z1h.pin_memory(), z2h.pin_memory()
    labels_h = labels_h.to(torch.int32).pin_memory()
    loss_h = torch.empty(1, dtype=torch.float32).pin_memory()

    mode = nat.MODE_SOFT
    crit = spcl_b200.SelfPacedSupConLoss(temperature=TAU, weight_update=MODE_NAME, precision="bf16",
                                         check_nan=False, validate=False)   # = the reference under `python -O`
    crit.set_gamma(GAMMA)

    def fwd_bwd(a, b, lab):
        if world == 1:
            loss = crit(a, b, target=lab)
        else:
            loss, _ = sharded_supcon_loss(a, b, lab, temperature=TAU, gamma=GAMMA, mode=mode)
        loss.backward()
        return loss

    a = z1h.to(dev).requires_grad_(True)
    b = z2h.to(dev).requires_grad_(True)
    lab = labels_h.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- device-resident timing (value) ----------------
    for _ in range(args.warmup):
        a.grad = b.grad = None
        fwd_bwd(a, b, lab)
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for s, e in evs:
        a.grad = b.grad = None
        flush.zero_()
        s.record()
        loss = fwd_bwd(a, b, lab)
        e.record()
    sync_all()
    clocks = sampler.stop()
    t_ms = sum(s.elapsed_time(e) for s, e in evs)
    if world > 1:
        t = torch.tensor([t_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms = t.item()
    ms_per_step = t_ms / args.steps
    value = N * N / (ms_per_step * 1e-3)
    loss_val = loss.item()

    # ---------------- dominant kernel alone: fused backward (roofline) ----------------
    rows = N // world
    peak, peak_sustained, peak_src = _peaks()
    if world == 1:
        scalars, row_stats, zpack, labels_full, sig = ops.supcon_fwd(a.detach(), b.detach(), lab, None, TAU, GAMMA,
                                                                     mode, False, True)
        gone = torch.ones(1, device=dev)
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(args.steps)]
        for _ in range(3):
            ops.supcon_bwd(gone, zpack, labels_full, sig, None, row_stats, scalars, TAU, GAMMA, mode, True, n_total, d)
        for s, e in kev:
            flush.zero_()
            s.record()
            ops.supcon_bwd(gone, zpack, labels_full, sig, None, row_stats, scalars, TAU, GAMMA, mode, True, n_total, d)
            e.record()
        torch.cuda.synchronize()
        bwd_ms = sum(s.elapsed_time(e) for s, e in kev) / args.steps
        fev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(args.steps)]
        for _ in range(3):                      # the custom-op dispatcher's first call carries one-time set-up
            ops.supcon_fwd(a.detach(), b.detach(), lab, None, TAU, GAMMA, mode, False, True)
        for s, e in fev:
            flush.zero_()
            s.record()
            ops.supcon_fwd(a.detach(), b.detach(), lab, None, TAU, GAMMA, mode, False, True)
            e.record()
        torch.cuda.synchronize()
        fwd_ms = sum(s.elapsed_time(e) for s, e in fev) / args.steps
        bwd_flops = 4.0 * rows * N * d
        traffic, traffic_src = _ncu_traffic("bwd")
        roofline = {"bound": "tensor", "kernel": "spcl::tc::bwd_kernel (+dZ memset)",
                    "achieved": bwd_flops / (bwd_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                    "frac": bwd_flops / (bwd_ms * 1e-3) / 1e12 / peak, "traffic": traffic,
                    "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                    "traffic_source": traffic_src,
                    "peak_source": f"{peak_src} burst (MEASURED_PEAKS.json bf16_tflops)",
                    "algorithmic_flops": bwd_flops, "kernel_ms": bwd_ms, "fwd_op_ms": fwd_ms,
                    "step_tflops_6N2d": 6.0 * rows * N * d / (ms_per_step * 1e-3) / 1e12,
                    "step_frac_of_peak": 6.0 * rows * N * d / (ms_per_step * 1e-3) / 1e12 / peak}
    else:
        step_tf = 6.0 * rows * N * d / (ms_per_step * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "whole step per GPU (fwd+bwd kernels + NCCL gathers)",
                    "achieved": step_tf, "peak": peak_sustained, "unit": "TFLOP/s", "frac": step_tf / peak_sustained,
                    "traffic": None, "peak_source": f"{peak_src} sustained"}

    # ---------------- end to end through the public API, host buffers ----------------
    # Every step copies its inputs from pinned host memory and reads its loss back.  (1) serial: copy, step,
    # read-back, synchronise -- the latency of ONE call; (2) streamed (the `e2e` value): spcl_b200.HostFeed stages
    # batch k + 1 on a side stream while batch k is in the kernels -- the throughput a host-fed caller gets.
    e2e_steps = args.steps
    def serial_step():
        a2 = z1h.to(dev, non_blocking=True).requires_grad_(True)
        b2 = z2h.to(dev, non_blocking=True).requires_grad_(True)
        lab2 = labels_h.to(dev, non_blocking=True)
        l2 = fwd_bwd(a2, b2, lab2)
        loss_h.copy_(l2.detach().reshape(1), non_blocking=True)
        torch.cuda.synchronize()
    for _ in range(3):                       # untimed: the caching allocator settles on this loop's block sizes
        serial_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        serial_step()
    if world > 1:
        dist.barrier()
    serial_s = (time.perf_counter() - t0) / e2e_steps

    feed = spcl_b200.HostFeed(n_loc, d, dev, depth=2)
    def streamed(k_steps):
        feed.push(z1h, z2h, labels_h)
        for k in range(k_steps):
            if k + 1 < k_steps:
                feed.push(z1h, z2h, labels_h)
            a2, b2, lab2, slot = feed.pop()
            l2 = fwd_bwd(a2, b2, lab2)
            feed.release(slot, l2)
        return feed.losses()                      # synchronises
    streamed(3)
    sync_all()
    feed.h2d_bytes = 0
    t0 = time.perf_counter()
    e2e_losses = streamed(e2e_steps)
    if world > 1:
        dist.barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    assert len(e2e_losses) == e2e_steps and abs(e2e_losses[-1] - loss_val) <= 1e-3 * abs(loss_val), e2e_losses[-3:]
    if world > 1:
        t = torch.tensor([e2e_s, serial_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, serial_s = t.tolist()
    # how much of the serial figure is the PCIe copy alone
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        z1h.to(dev, non_blocking=True)
        z2h.to(dev, non_blocking=True)
        torch.cuda.synchronize()
    h2d_ms = (time.perf_counter() - t0) / 5 * 1e3
    e2e = {"value": N * N / e2e_s, "unit": "pairs/s", "ms_per_step": e2e_s * 1e3,
           "how": "spcl_b200.HostFeed: pinned host -> 2 staging slots on a copy stream, fused fwd+bwd, loss -> pinned host; every step copies and computes",
           "serial_ms_per_step": serial_s * 1e3, "serial_value": N * N / serial_s, "h2d_only_ms": h2d_ms,
           "h2d_bytes_per_step": feed.h2d_bytes // e2e_steps, "d2h_bytes_per_step": 4}

    # ---------------- the reference's own batch size (cfg2), secondary figure ----------------
    # K = 3 meta-label problems of 2 x 256 anchors, d = 256 (encoder pre-training) through the public
    # modules: fwd + bwd of all three, eager launches vs cuda_graph=True (one graph replay per loss call).
    small = None
    if world == 1:
        from spcl_b200.workloads import acdc_meta_labels, make_views
        meta = acdc_meta_labels(256)
        def small_step(graphed, grouped=False, group_graph=False):
            probs = []
            for kind, gm in (("partition", 5.0), ("patient", 3.5), ("cycle", 2.0)):
                c = spcl_b200.SelfPacedSupConLoss(weight_update="soft", correct_grad=True, check_nan=False,
                                                  validate=False, cuda_graph=graphed)
                c.set_gamma(gm)
                v1, v2 = make_views(meta[kind], 256, sigma=0.7, seed=1)
                probs.append((c, v1.to(dev).requires_grad_(True), v2.to(dev).requires_grad_(True),
                              meta[kind].int().to(dev)))
            def run_grouped():
                for _, p1, p2, _lb in probs:
                    p1.grad = p2.grad = None
                ls = spcl_b200.grouped_forward([q[0] for q in probs], [(q[1], q[2]) for q in probs],
                                               [q[3] for q in probs], cuda_graph=group_graph)
                (ls[0] + ls[1] + ls[2]).backward()
            def run():
                if grouped:
                    return run_grouped()
                tot = None
                for c, p1, p2, lb in probs:
                    p1.grad = p2.grad = None
                    l = c(p1, p2, target=lb)
                    tot = l if tot is None else tot + l
                tot.backward()
            for _ in range(10):
                run()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(100):
                run()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / 100 * 1e6
        small = {"workload": "cfg2: 3 meta-label problems, N = 512, d = 256, fp32 path, fwd+bwd of all three",
                 "eager_us_per_step": small_step(False), "cuda_graph_us_per_step": small_step(True),
                 "grouped_launch_us_per_step": small_step(False, grouped=True),
                 "grouped_graph_us_per_step": small_step(False, grouped=True, group_graph=True)}

    # ---------------- dense front end, secondary figure: HBM-bound kernels ----------------
    dense_fe = None
    if world == 1:
        import importlib.util
        mod_spec = importlib.util.spec_from_file_location("gpu_dense_bench", os.path.join(os.path.dirname(
            os.path.abspath(__file__)), "tools", "gpu_dense_bench.py"))
        try:                                        # a secondary figure must not take the headline down with it
            gdb = importlib.util.module_from_spec(mod_spec)
            mod_spec.loader.exec_module(gdb)
            r = gdb.measure(cases=gdb.CASES[:1], reps=10)
            c = r["cases"][0]
            dense_fe = {"workload": "DenseProjectionHead tail: 32 x 128 x 224 x 224 -> 32 x 32 pooled, normalised rows",
                        "bound": "hbm", "peak": r["peak_gbs"], "unit": "GB/s", "fwd_ms": c["fwd_ms"], "bwd_ms": c["bwd_ms"],
                        "fwd_achieved": c["fwd_gbs"], "fwd_frac": c["fwd_frac"], "bwd_achieved": c["bwd_gbs"],
                        "bwd_frac": c["bwd_frac"], "l2": c["l2"]}
        except Exception as e:
            dense_fe = {"error": f"{type(e).__name__}: {e}"}

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    cpu = None
    if world == 1 and not args.skip_cpu_baseline:
        r = cpu_port_time(workload, 3, 1)
        cpu = {"value": r["pairs_per_s"], "unit": "pairs/s", "cores": r["cores"], "kind": "port",
               "sample": r["sample"], "ms_per_step_sample": r["ms"]}

    if rank == 0:
        line = {
            "metric": "supcon_fwd_bwd_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload, "anchors_N": N, "d": d, "rows_per_gpu": rows,
                       "hyper": {"tau": TAU, "gamma": GAMMA, "mode": MODE_NAME, "labels": spec["labels"]},
                       "l2": "256 MB flush between timed steps", "parallelism": f"row-shard x{world}"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "small_batch": small, "dense_front_end": dense_fe, "gpu_launches": 6 * args.steps,
            "clocks": clocks, "loss": loss_val,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
