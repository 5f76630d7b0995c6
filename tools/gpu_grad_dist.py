"""Development tool: per-element distribution of the bf16 path's gradient error against the fp64 reference at cfg3
(the numbers behind the percentile bounds in tests/test_gpu_parity.py::test_cfg3_full_size_against_fp64)."""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

import spcl_b200  # noqa: E402
from spcl_b200.workloads import make_workload  # noqa: E402
from torch_ref import supcon_ref64  # noqa: E402

for workload, mode, gamma in (("cfg3_dense_2x16384_d128_simclr", "soft", 8.0), ("cfg3_dense_2x16384_d128_slice", "soft", 10.0),
                              ("cfg3_dense_2x16384_d128_slice", "hard", 8.5)):
    z1, z2, labels = make_workload(workload)
    z1, z2 = z1.bfloat16().float().cuda(), z2.bfloat16().float().cuda()
    lab = labels.int().cuda()
    crit = spcl_b200.SelfPacedSupConLoss(weight_update=mode, precision="bf16", validate=False)
    crit.set_gamma(gamma)
    a, b = z1.clone().requires_grad_(True), z2.clone().requires_grad_(True)
    crit(a, b, target=lab).backward()
    ref = supcon_ref64(z1, z2, lab, gamma=gamma, mode=mode)
    g = torch.cat([a.grad, b.grad]).double()
    r = torch.cat([ref["dz1"], ref["dz2"]])
    err = (g - r).abs() / r.abs().max()
    q = torch.quantile(err.flatten()[:: 7].float(), torch.tensor([0.5, 0.99, 0.999], device=err.device))
    rows = (g - r).norm(dim=1) / r.norm(dim=1).clamp_min(1e-300)
    print(f"{workload} {mode} g={gamma}: max {err.max().item():.2e} p50 {q[0].item():.2e} p99 {q[1].item():.2e} p99.9 {q[2].item():.2e} "
          f"rms {err.pow(2).mean().sqrt().item():.2e} | per-row rel L2: max {rows.max().item():.2e} median {rows.median().item():.2e}",
          flush=True)
