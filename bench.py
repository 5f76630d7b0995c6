#!/usr/bin/env python
"""Headline benchmark: self-paced SupCon fwd+bwd pairs/sec (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one forward + backward of the self-paced loss over one synthetic batch.
  N = 1 : cfg3  dense pixel contrast, 2 x 16384 anchors, d = 128  (the config the metric is quoted on)
  N > 1 : cfg4  N = 262144 anchors, d = 128, anchor rows sharded over the ranks (torchrun, NCCL)
Prints ONE JSON line (rank 0).  `value` times the device-resident step with CUDA events; `e2e` times the
public module API from pinned host buffers (H2D + step + D2H of the loss); `roofline` is the backward
kernel alone against the measured bf16 peak; `cpu_baseline` is the UNMODIFIED reference module
(baseline/_ref/contrastyou/losses/contrast_loss3.py, installed by tools/install_ref.sh) on the host cores
(`kind: "reference"`; the oracle's dense port only if that directory is missing).  `--impl reference` times
that same CPU module as the reference arm.

Extra keys of the N = 1 line: `extra` (cfg3 with slice labels -- dense positives -- soft and hard weighting),
`gpu_eager_reference` (the unmodified reference module on this GPU, largest N that completes),
`aux_hbm` (HBM fractions of the normalise / pack kernels), `cfg5` (encoder step on the reference UNet).
N > 1: `parity` (every rank checks its slab of the sharded result against a blockwise fp64 reference; the
run FAILS outside the stated tolerances), `strong_scaling_base_ms` (the same cfg4 problem on ONE GPU),
`efficiency_vs_cfg4_1gpu`, `per_rank` clocks.
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
CPU_SAMPLE_N = 4096          # anchors per view of the bounded CPU sample (N = 8192: ~7 GB of N x N fp32 temporaries)


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


def cpu_reference_time(workload: str, steps: int, warmup: int):
    """Times the reference's own CPU implementation of the path on the host cores: the unmodified
    ``SelfPacedSupConLoss`` of contrast_loss3.py from baseline/_ref (``kind = "reference"``); if that directory is
    missing, the oracle's dense fp32 port of the same op sequence (``kind = "port"``)."""
    import torch
    from spcl_b200.workloads import make_workload
    torch.set_num_threads(os.cpu_count() or 1)
    kind, note = "reference", "unmodified baseline/_ref/contrastyou/losses/contrast_loss3.py (SelfPacedSupConLoss)"
    crit = None
    try:
        from baseline import ref_loader
        mod = ref_loader.loss_module()
        crit = mod.SelfPacedSupConLoss(temperature=TAU, weight_update=MODE_NAME, correct_grad=False)
        crit.set_gamma(GAMMA)
    except Exception as e:                                  # noqa: BLE001
        kind, note = "port", f"oracle/dense_port.py (reference files unavailable: {type(e).__name__}: {e})"
    z1, z2, labels = make_workload(workload)
    n = CPU_SAMPLE_N
    z1, z2, labels = z1[:n].clone(), z2[:n].clone(), labels[:n].tolist()
    times = []
    for i in range(warmup + steps):
        a = z1.clone().requires_grad_(True)
        b = z2.clone().requires_grad_(True)
        t0 = time.perf_counter()
        if crit is not None:
            loss = crit(a, b, target=labels)
        else:
            from oracle.dense_port import dense_supcon      # the CPU baseline leg may execute oracle/
            loss = dense_supcon(a, b, target=labels, temperature=TAU, gamma=GAMMA, mode=MODE_NAME).loss
        loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    N = 2 * n
    return dict(ms=1e3 * statistics.mean(times), best_ms=1e3 * min(times), N=N, cores=torch.get_num_threads(),
                pairs_per_s=N * N / statistics.mean(times), kind=kind, loss=float(loss.item()),
                sample=f"first {n} anchors/view of {workload} (N={N}, d={z1.shape[1]}), fp32, forward + autograd "
                       f"backward of {note}; the full N=32768 needs ~26 live N x N fp32 matrices (110 GB) on the host")


def run_reference(args, workload, world, rank):
    if rank != 0:
        return
    r = cpu_reference_time(workload, args.steps, max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": "supcon_fwd_bwd_pairs_per_sec", "value": r["pairs_per_s"], "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "hyper": {"tau": TAU, "gamma": GAMMA, "mode": MODE_NAME},
                   "note": "the reference is CPU PyTorch; each step is a bounded sample of the workload (see "
                           "cpu_baseline.sample): same_config differs from the GPU arm in N only",
                   "sample_anchors_N": r["N"]},
        "cpu_baseline": {"value": r["pairs_per_s"], "unit": "pairs/s", "cores": r["cores"], "kind": r["kind"],
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


def timed_steps(fn, steps, flush):
    """ms per call of fn, device-timed with an L2 flush before every call."""
    import torch
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for s, e in evs:
        flush.zero_()
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    return sum(s.elapsed_time(e) for s, e in evs) / steps


def extra_workloads(dev, flush, steps, peak):
    """cfg3 (b) of SURVEY 8d: slice labels (1024 pixels share a label: dense positives), soft and hard weighting."""
    import torch
    import spcl_b200
    from spcl_b200.workloads import WORKLOADS, make_workload
    out = []
    name = "cfg3_dense_2x16384_d128_slice"
    z1h, z2h, labels_h = make_workload(name, seed=0)
    a = z1h.to(dev).requires_grad_(True)
    b = z2h.to(dev).requires_grad_(True)
    lab = labels_h.to(torch.int32).to(dev)
    n, d = z1h.shape
    N = 2 * n
    for mode_name, gamma in (("soft", 10.0), ("hard", 10.0)):
        crit = spcl_b200.SelfPacedSupConLoss(temperature=TAU, weight_update=mode_name, precision="bf16",
                                             check_nan=False, validate=False)
        crit.set_gamma(gamma)

        def step():
            a.grad = b.grad = None
            loss = crit(a, b, target=lab)
            loss.backward()
            return loss
        for _ in range(3):
            loss = step()
        ms = timed_steps(step, steps, flush)
        tf = 6.0 * N * N * d / (ms * 1e-3) / 1e12
        out.append({"workload": name, "labels": WORKLOADS[name]["labels"], "mode": mode_name, "gamma": gamma,
                    "ms_per_step": ms, "pairs_per_s": N * N / (ms * 1e-3), "tflops_6N2d": tf, "frac_of_peak": tf / peak,
                    "loss": float(loss.item()), "downgrade_ratio": float(crit.downgrade_ratio)})
    return out


def gpu_eager_reference(dev, z1h, z2h, flush, ours_ms_full):
    """The UNMODIFIED reference module (baseline/_ref contrast_loss3.SelfPacedSupConLoss) on this GPU, PyTorch eager,
    same inputs and hyper-parameters, at the largest N of {32768, 16384, 8192} that completes (every N x N fp32
    temporary is 4.3 GB at N = 32768 and the module keeps ~26 of them alive)."""
    import torch
    import spcl_b200
    try:
        from baseline import ref_loader
        mod = ref_loader.loss_module()
    except Exception as e:                                  # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"}
    tried = []
    for n in (z1h.shape[0], z1h.shape[0] // 2, z1h.shape[0] // 4):
        crit = mod.SelfPacedSupConLoss(temperature=TAU, weight_update=MODE_NAME, correct_grad=False)
        crit.set_gamma(GAMMA)
        labels = list(range(n))
        try:
            torch.cuda.reset_peak_memory_stats(dev)
            a = z1h[:n].to(dev).requires_grad_(True)
            b = z2h[:n].to(dev).requires_grad_(True)

            def step():
                a.grad = b.grad = None
                loss = crit(a, b, target=labels)
                loss.backward()
                return loss
            loss = step()                                       # warm-up (allocator, cuBLAS)
            torch.cuda.synchronize()
            ms = timed_steps(step, 3, flush)
            peak_gb = torch.cuda.max_memory_allocated(dev) / 1e9
            ref_loss = float(loss.item())
            # drop the N x N diagnostics the module keeps (sim_exp, sim_logits, masks)
            for k in ("sim_exp", "sim_logits", "pos_mask", "neg_mask", "sp_mask"):
                if hasattr(crit, k):
                    delattr(crit, k)
            del loss
            N = 2 * n
            res = {"impl": "baseline/_ref/contrastyou/losses/contrast_loss3.py SelfPacedSupConLoss, torch eager, fp32",
                   "anchors_N": N, "ms_per_step": ms, "pairs_per_s": N * N / (ms * 1e-3), "peak_mem_gb": peak_gb,
                   "loss": ref_loss, "oom_at": tried}
            if n == z1h.shape[0]:
                res["ours_ms_same_N"] = ours_ms_full
            else:                                               # measure this repo at the same reduced N
                c2 = spcl_b200.SelfPacedSupConLoss(temperature=TAU, weight_update=MODE_NAME, precision="bf16",
                                                   check_nan=False, validate=False)
                c2.set_gamma(GAMMA)
                lab = torch.arange(n, dtype=torch.int32, device=dev)

                def ours():
                    a.grad = b.grad = None
                    c2(a, b, target=lab).backward()
                for _ in range(3):
                    ours()
                res["ours_ms_same_N"] = timed_steps(ours, 10, flush)
            res["speedup_same_N"] = res["ms_per_step"] / res["ours_ms_same_N"]
            return res
        except torch.cuda.OutOfMemoryError:
            tried.append(2 * n)
        finally:
            crit = None
            a = b = None
            import gc
            gc.collect()
            torch.cuda.empty_cache()
    return {"unavailable": "out of memory at every tried N", "oom_at": tried}


def _load_tool(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, str(ROOT / "tools" / f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def sharded_parity(dist, dev, a, b, lab, world, rank, loss_val, scalars):
    """Every rank checks ITS slab of the sharded result -- loss, ratio, logD and the gradient rows of its own anchors --
    against a blockwise fp64 reference (tests/torch_ref.py, itself pinned to the numpy oracle) computed on the gathered
    bf16-rounded operands.  -> dict (max over ranks) with `ok`."""
    import torch
    sys.path.insert(0, str(ROOT / "tests"))
    from torch_ref import supcon_ref64_rows
    n_loc, d = a.shape
    rows_loc = 2 * n_loc
    N = rows_loc * world
    # the operands the kernels saw: bf16-rounded rows in the global (rank, view, sample) order
    z_loc = torch.cat([a.detach(), b.detach()]).bfloat16()
    z_all = torch.empty(N, d, dtype=torch.bfloat16, device=dev)
    dist.all_gather_into_tensor(z_all, z_loc)
    lab_all = torch.empty(N, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(lab_all, torch.cat([lab, lab]))

    def reduce_sum(t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def gather_rows(t):                                         # [k, rows_loc] per rank -> [k, N]
        out = torch.empty(world, *t.shape, dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(out, t.contiguous())
        return out.permute(1, 0, 2).reshape(t.shape[0], -1)

    t0 = time.perf_counter()
    ref = supcon_ref64_rows(z_all.float(), lab_all, rank * rows_loc, (rank + 1) * rows_loc, temperature=TAU,
                            gamma=GAMMA, mode=MODE_NAME, reduce_sum=reduce_sum, gather_rows=gather_rows)
    torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    g = torch.cat([a.grad, b.grad]).double()
    r = ref["dZ"]
    gmax = r.abs().max()
    cos = (g * r).sum() / (g.norm() * r.norm())
    local = torch.tensor([abs(loss_val - ref["loss"]) / abs(ref["loss"]),
                          abs(float(scalars[1]) - ref["ratio"]) / abs(ref["ratio"]),
                          ((g - r).abs().max() / gmax).item(), 1.0 - cos.item()], dtype=torch.float64, device=dev)
    dist.all_reduce(local, op=dist.ReduceOp.MAX)
    loss_rel, ratio_rel, grad_max_rel, one_minus_cos = local.tolist()
    tol = {"loss_rel": 3e-4, "ratio_rel": 2e-4, "grad_max_rel": 3e-2, "grad_cos_min": 0.9999}
    ok = (loss_rel <= tol["loss_rel"] and ratio_rel <= tol["ratio_rel"] and grad_max_rel <= tol["grad_max_rel"]
          and 1.0 - one_minus_cos >= tol["grad_cos_min"])
    return {"ok": bool(ok), "loss_rel": loss_rel, "ratio_rel": ratio_rel, "grad_max_rel": grad_max_rel,
            "grad_cos": 1.0 - one_minus_cos, "tolerances": tol, "reference": "tests/torch_ref.py supcon_ref64_rows "
            "(fp64, blockwise) on the gathered bf16-rounded operands; each rank checks its own rows against ALL columns; "
            "figures are the max over ranks", "rows_checked_per_rank": rows_loc, "seconds": secs,
            "loss_ref": ref["loss"], "loss": loss_val}


def cfg4_single_gpu_ms(dev, world, n_loc, d, flush, steps=3):
    """The SAME cfg4 problem (all world * n_loc anchors per view) on ONE GPU through the single-GPU module: the base of
    the strong-scaling curve.  Run on rank 0 only."""
    import torch
    import spcl_b200
    zs1, zs2 = [], []
    for r in range(world):
        g = torch.Generator().manual_seed(1000 + r)
        base = torch.randn(n_loc, d, generator=g)
        zs1.append(torch.nn.functional.normalize(base + 0.7 * torch.randn(n_loc, d, generator=g), dim=1))
        zs2.append(torch.nn.functional.normalize(base + 0.7 * torch.randn(n_loc, d, generator=g), dim=1))
    a = torch.cat(zs1).to(dev).requires_grad_(True)
    b = torch.cat(zs2).to(dev).requires_grad_(True)
    lab = torch.arange(world * n_loc, dtype=torch.int32, device=dev)
    crit = spcl_b200.SelfPacedSupConLoss(temperature=TAU, weight_update=MODE_NAME, precision="bf16", check_nan=False,
                                         validate=False)
    crit.set_gamma(GAMMA)

    def step():
        a.grad = b.grad = None
        loss = crit(a, b, target=lab)
        loss.backward()
        return loss
    for _ in range(2):
        loss = step()
    ms = timed_steps(step, steps, flush)
    return ms, float(loss.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline + roofline + e2e only (profiling runs)")
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

    last = {}

    def fwd_bwd(a, b, lab):
        if world == 1:
            loss = crit(a, b, target=lab)
        else:
            loss, last["scalars"] = sharded_supcon_loss(a, b, lab, temperature=TAU, gamma=GAMMA, mode=mode)
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
    # everything host-side that differs in duration from rank to rank (NVML start-up, event creation) happens BEFORE the
    # barrier: a rank that enters the first timed step late makes every other rank wait inside its timed region
    # (r02s on 8 GPUs: first step 7-18 ms against 5.75 ms for the following ones)
    sampler = ClockSampler(local_rank)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler.start()
    sync_all()
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

    # ---------------- multi-GPU: parity of the sharded result, strong-scaling base, per-rank clocks ----------------
    parity = base_ms = per_rank = None
    if world > 1:
        parity = sharded_parity(dist, dev, a, b, lab, world, rank, loss_val, last["scalars"].tolist())
        gathered = [None] * world
        dist.all_gather_object(gathered, {"rank": rank, "ms_per_step_local": sum(s.elapsed_time(e) for s, e in evs) / args.steps,
                                          "step_ms": [round(s.elapsed_time(e), 3) for s, e in evs[:12]], **clocks})
        per_rank = gathered
        base = torch.zeros(2, dtype=torch.float64, device=dev)
        if rank == 0 and not args.quick:
            b_ms, b_loss = cfg4_single_gpu_ms(dev, world, n_loc, d, flush)
            base[0], base[1] = b_ms, b_loss
        dist.broadcast(base, src=0)
        base_ms = float(base[0]) or None
        if base_ms is not None and abs(float(base[1]) - loss_val) > 3e-4 * abs(loss_val):
            parity["ok"] = False
            parity["single_gpu_loss_mismatch"] = [float(base[1]), loss_val]
        torch.cuda.empty_cache()

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
           "how": "spcl_b200.HostFeed: pinned host -> 2 staging slots on a copy stream, fused fwd+bwd, loss -> pinned host; "
                  "every step copies and computes.  Only the 4-byte loss returns to the host: the gradients (16.8 MB at "
                  "cfg3) stay on the device, where the caller's projector / backbone backward consumes them",
           "serial_ms_per_step": serial_s * 1e3, "serial_value": N * N / serial_s, "h2d_only_ms": h2d_ms,
           "h2d_bytes_per_step": feed.h2d_bytes // e2e_steps, "d2h_bytes_per_step": 4}

    # ---------------- the reference's own batch size (cfg2), secondary figure ----------------
    # K = 3 meta-label problems of 2 x 256 anchors, d = 256 (encoder pre-training) through the public
    # modules: fwd + bwd of all three, eager launches vs cuda_graph=True (one graph replay per loss call).
    small = None
    if world == 1 and not args.quick:
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
                 "route": "one cooperative launch per loss call / per group (spcl_supcon_group_fused_f32): S kept in "
                          "registers across the stages, backward = one multiply",
                 "eager_us_per_step": small_step(False), "cuda_graph_us_per_step": small_step(True),
                 "grouped_launch_us_per_step": small_step(False, grouped=True),
                 "grouped_graph_us_per_step": small_step(False, grouped=True, group_graph=True)}
        try:                                        # device time of the group through the C ABI alone
            import importlib.util
            gs_spec = importlib.util.spec_from_file_location("gpu_small", os.path.join(os.path.dirname(
                os.path.abspath(__file__)), "tools", "gpu_small.py"))
            gs = importlib.util.module_from_spec(gs_spec)
            gs_spec.loader.exec_module(gs)
            small["c_abi_device_us"] = gs.raw_group_times(256, 256)
        except Exception as e:                      # noqa: BLE001
            small["c_abi_device_us"] = {"error": f"{type(e).__name__}: {e}"}
        os.environ["SPCL_FUSED_SMALL"] = "0"        # the staged route (one launch per stage), same harness
        try:
            small["staged_route"] = {"eager_us_per_step": small_step(False), "cuda_graph_us_per_step": small_step(True),
                                     "grouped_launch_us_per_step": small_step(False, grouped=True)}
        finally:
            del os.environ["SPCL_FUSED_SMALL"]

    # ---------------- dense front end, secondary figure: HBM-bound kernels ----------------
    dense_fe = None
    if world == 1 and not args.quick:
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
                        "bwd_frac": c["bwd_frac"], "l2": c["l2"], "fwd_queued_ms": c["fwd_queued_ms"],
                        "bwd_queued_ms": c["bwd_queued_ms"], "fwd_queued_frac": c["fwd_queued_frac"],
                        "bwd_queued_frac": c["bwd_queued_frac"], "queued": c["queued"]}
        except Exception as e:
            dense_fe = {"error": f"{type(e).__name__}: {e}"}

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    cpu = None
    if world == 1 and not args.skip_cpu_baseline and not args.quick:
        r = cpu_reference_time(workload, 3, 1)
        cpu = {"value": r["pairs_per_s"], "unit": "pairs/s", "cores": r["cores"], "kind": r["kind"],
               "sample": r["sample"], "ms_per_step_sample": r["ms"]}

    # ---------------- secondary figures (N = 1): dense-positive labelling, eager reference on this GPU, aux kernels, cfg5
    extra = eager = aux_hbm = cfg5 = None
    if world == 1 and not args.quick:
        def guarded(fn):                            # a secondary figure must not take the headline down with it
            try:
                return fn()
            except Exception as e:                  # noqa: BLE001
                torch.cuda.empty_cache()
                return {"error": f"{type(e).__name__}: {e}"}
        extra = guarded(lambda: extra_workloads(dev, flush, args.steps, peak))
        aux_hbm = guarded(lambda: _load_tool("gpu_aux_bench").measure())
        cfg5 = guarded(lambda: _load_tool("cfg5_step").measure(batch=64, steps=5, warmup=3))
        del a, b
        torch.cuda.empty_cache()
        eager = guarded(lambda: gpu_eager_reference(dev, z1h, z2h, flush, ms_per_step))

    if world > 1 and parity is not None and not parity["ok"]:
        if rank == 0:
            print(json.dumps({"error": "sharded result outside the stated tolerances", "parity": parity}), flush=True)
        dist.destroy_process_group()
        sys.exit(3)

    if rank == 0:
        line = {
            "metric": "supcon_fwd_bwd_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload, "anchors_N": N, "d": d, "rows_per_gpu": rows,
                       "hyper": {"tau": TAU, "gamma": GAMMA, "mode": MODE_NAME, "labels": spec["labels"]},
                       "l2": "256 MB flush between timed steps", "parallelism": f"row-shard x{world}"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "small_batch": small, "dense_front_end": dense_fe,
            "gpu_launches": 7 * args.steps,     # prepare, stats, sp, row_finalize, finalize, transpose, bwd (+ 2 memsets)
            "clocks": clocks, "loss": loss_val, "extra": extra, "gpu_eager_reference": eager, "aux_hbm": aux_hbm,
            "cfg5": cfg5,
        }
        if world > 1:
            line["parity"] = parity
            line["per_rank"] = per_rank
            line["strong_scaling_base_ms"] = base_ms
            line["strong_scaling_base"] = "the same cfg4 problem (N = 262144) on ONE GPU, single-GPU module, device-timed"
            line["efficiency_vs_cfg4_1gpu"] = (base_ms / (world * ms_per_step)) if base_ms else None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
