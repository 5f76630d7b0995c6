"""GPU bring-up diagnostics (development tool, not part of the product or the test-suite).

Runs each stage in a fresh subprocess (a device trap must not poison the next stage) and prints
numeric error summaries against the oracle and between the fp32 SIMT and bf16 tensor-core paths.
    python tools/gpu_diag.py [stage ...]
"""
import json
import os
import pathlib
import subprocess
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _inputs(n, d, kind, seed=0):
    import torch
    from spcl_b200.workloads import acdc_meta_labels, make_views
    if kind == "self":
        labels = torch.arange(n)
    elif kind == "slice":
        labels = torch.arange(n) // max(1, n // 16)
    else:
        labels = acdc_meta_labels(n)[kind]
    z1, z2 = make_views(labels, d, sigma=0.7, seed=seed)
    return z1, z2, labels


def stage_simt(n=64, d=128):
    import numpy as np, torch
    from oracle.closed_form import supcon_closed_form
    import spcl_b200 as sp
    z1, z2, labels = _inputs(n, d, "partition")
    for mode, name in ((0, "none"), (1, "hard"), (2, "soft")):
        a = z1.cuda().requires_grad_(True); b = z2.cuda().requires_grad_(True)
        loss, sc, aux = sp.supcon_loss(a, b, target=labels.tolist(), gamma=5.0, mode=mode, precision="fp32")
        loss.backward()
        ref = supcon_closed_form(z1.numpy(), z2.numpy(), target=labels.tolist(), gamma=5.0, mode=name)
        ge = np.abs(a.grad.cpu().numpy() - ref["dz1"]).max() / np.abs(ref["dz1"]).max()
        print(f"simt {name}: loss {loss.item():.7f} ref {ref['loss']:.7f} ratio {sc[1].item():.6f}/{ref['ratio']:.6f} grad relerr {ge:.2e}")


def stage_tc_fwd(n=256, d=128, kind="partition"):
    import numpy as np, torch
    from oracle.closed_form import supcon_closed_form
    import spcl_b200 as sp
    z1, z2, labels = _inputs(n, d, kind)
    zb1, zb2 = z1.bfloat16().float(), z2.bfloat16().float()
    for mode, name in ((0, "none"), (1, "hard"), (2, "soft")):
        loss, sc, aux = sp.supcon_loss(z1.cuda(), z2.cuda(), target=labels.tolist(), gamma=5.0, mode=mode, precision="bf16")
        ref = supcon_closed_form(zb1.numpy(), zb2.numpy(), target=labels.tolist(), gamma=5.0, mode=name, want_grad=False)
        st = aux["row_stats"][:, :2 * n].cpu().numpy().T
        print(f"tc fwd n={n} d={d} {kind} {name}: loss {loss.item():.6f} ref(bf16 in) {ref['loss']:.6f} ratio {sc[1].item():.5f}/{ref['ratio']:.5f} "
              f"logD err {np.abs(st[:,0]-ref['logD']).max():.2e} c err {np.abs(1/st[:,1]-ref['c']).max():.2e}")


def stage_tc_bwd(n=256, d=128, kind="partition"):
    import numpy as np, torch
    from oracle.closed_form import supcon_closed_form
    import spcl_b200 as sp
    z1, z2, labels = _inputs(n, d, kind)
    zb1, zb2 = z1.bfloat16().float(), z2.bfloat16().float()
    for mode, name in ((0, "none"), (2, "soft"), (1, "hard")):
        a = z1.cuda().requires_grad_(True); b = z2.cuda().requires_grad_(True)
        loss, sc, aux = sp.supcon_loss(a, b, target=labels.tolist(), gamma=5.0, mode=mode, precision="bf16")
        loss.backward()
        ref = supcon_closed_form(zb1.numpy(), zb2.numpy(), target=labels.tolist(), gamma=5.0, mode=name)
        g = np.concatenate([a.grad.cpu().numpy(), b.grad.cpu().numpy()]); r = np.concatenate([ref["dz1"], ref["dz2"]])
        cos = (g * r).sum() / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-30)
        print(f"tc bwd n={n} d={d} {kind} {name}: loss {loss.item():.6f}/{ref['loss']:.6f} grad max relerr {np.abs(g-r).max()/np.abs(r).max():.3e} cos {cos:.6f} |g| {np.linalg.norm(g):.4e} |r| {np.linalg.norm(r):.4e}")


def stage_tc_vs_simt(n=2048, d=128, kind="slice"):
    import torch
    import spcl_b200 as sp
    z1, z2, labels = _inputs(n, d, kind)
    z1 = z1.bfloat16().float(); z2 = z2.bfloat16().float()      # same operand values on both paths
    out = {}
    for prec in ("fp32", "bf16"):
        a = z1.cuda().requires_grad_(True); b = z2.cuda().requires_grad_(True)
        loss, sc, aux = sp.supcon_loss(a, b, target=labels.int().cuda(), gamma=8.0, mode=2, precision=prec)
        loss.backward()
        out[prec] = (loss.item(), sc[1].item(), torch.cat([a.grad, b.grad]))
    g32, g16 = out["fp32"][2], out["bf16"][2]
    print(f"tc vs simt n={n} d={d} {kind}: loss {out['bf16'][0]:.6f} vs {out['fp32'][0]:.6f}; ratio {out['bf16'][1]:.5f} vs {out['fp32'][1]:.5f}; "
          f"grad relerr {(g16-g32).abs().max().item()/g32.abs().max().item():.3e} cos {torch.nn.functional.cosine_similarity(g16.flatten(), g32.flatten(), dim=0).item():.6f}")


def stage_time(n=16384, d=128, kind="self", prec="bf16", mode=2):
    import torch
    import spcl_b200 as sp
    z1, z2, labels = _inputs(n, d, kind)
    a = z1.cuda().requires_grad_(True); b = z2.cuda().requires_grad_(True)
    lab = labels.int().cuda()
    def step():
        a.grad = None; b.grad = None
        loss, sc, aux = sp.supcon_loss(a, b, target=lab, gamma=8.0, mode=mode, precision=prec)
        loss.backward()
        return loss
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    e0.record()
    for _ in range(K): l = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    N = 2 * n
    print(f"time n={n} d={d} {kind} {prec} mode={mode}: {ms:.3f} ms/step  pairs/s {N*N/ms*1e3:.3e}  TF/s(6N^2d) {6*N*N*d/ms*1e3/1e12:.1f} loss {l.item():.5f}")


def stage_l2norm():
    import torch, torch.nn.functional as F
    import spcl_b200 as sp
    for shape, dim in (((4096, 256), 1), ((32, 128, 32, 32), 1), ((7, 33), 1), ((3, 5, 7), 1)):
        x = torch.randn(*shape, device="cuda", requires_grad=True)
        y = sp.normalize(x, dim=dim); g = torch.randn_like(y); y.backward(g)
        x2 = x.detach().clone().requires_grad_(True); y2 = F.normalize(x2, dim=dim); y2.backward(g)
        print(f"l2norm {shape}: fwd err {(y-y2).abs().max().item():.2e} bwd err {(x.grad-x2.grad).abs().max().item():.2e}")


STAGES = {
    "simt": "stage_simt()", "l2norm": "stage_l2norm()",
    "tc_fwd": "stage_tc_fwd()", "tc_fwd_self": "stage_tc_fwd(256,128,'self')", "tc_fwd_d256": "stage_tc_fwd(256,256,'composite')",
    "tc_fwd_ragged": "stage_tc_fwd(75,96,'partition')",
    "tc_bwd": "stage_tc_bwd()", "tc_bwd_self": "stage_tc_bwd(256,128,'self')", "tc_bwd_d256": "stage_tc_bwd(256,256,'composite')",
    "tc_bwd_ragged": "stage_tc_bwd(75,96,'partition')", "tc_bwd_d64": "stage_tc_bwd(192,64,'patient')",
    "tc_vs_simt": "stage_tc_vs_simt()", "tc_vs_simt_self": "stage_tc_vs_simt(4096,128,'self')",
    "time_bf16": "stage_time()", "time_bf16_none": "stage_time(mode=0)", "time_bf16_slice": "stage_time(kind='slice')",
    "time_fp32_4k": "stage_time(n=2048, prec='fp32')", "time_small": "stage_time(n=256, d=256, kind='composite')",
    "time_small_fp32": "stage_time(n=256, d=256, kind='composite', prec='fp32')",
}

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--run":
        eval(STAGES[sys.argv[2]])
        sys.exit(0)
    names = sys.argv[1:] or list(STAGES)
    for name in names:
        t0 = time.time()
        env = dict(os.environ)
        try:
            r = subprocess.run([sys.executable, __file__, "--run", name], capture_output=True, text=True, timeout=240, env=env)
            out = (r.stdout + ("\n[stderr] " + r.stderr[-1500:] if r.returncode != 0 else "")).strip()
            print(f"=== {name} (rc={r.returncode}, {time.time()-t0:.0f}s) ===\n{out}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"=== {name}: TIMEOUT ===", flush=True)
