"""Workload of tools/sanitize.sh: one small pass over every kernel family under compute-sanitizer.

smoke() (fp32 SIMT + bf16 tensor-core fwd/bwd, n = 192), a 1024-anchor bf16 case with slice labels in hard and soft
mode (tcgen05 / TMA / mbarrier pipelines of stats_kernel, sp_kernel, bwd_kernel, transpose_kernel), the grouped fp32
launch (the one-launch cooperative kernel and the staged route), the opt-in backward kernels, l2norm rows / NCHW fwd + bwd, the fused projector tail and the dense front end (pool rows / points)."""
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402
import spcl_b200  # noqa: E402
from spcl_b200 import ops  # noqa: E402
from spcl_b200.workloads import acdc_meta_labels, make_views  # noqa: E402


def main():
    g.smoke()
    n, d = 512, 128                                   # N = 1024 anchors
    labels = (torch.arange(n) // 32).int()
    z1, z2 = make_views(labels, d, sigma=0.7, seed=1)
    for mode, gamma in (("soft", 6.0), ("hard", 5.0)):
        crit = spcl_b200.SelfPacedSupConLoss(weight_update=mode, correct_grad=True, precision="bf16")
        crit.set_gamma(gamma)
        a, b = z1.cuda().requires_grad_(True), z2.cuda().requires_grad_(True)
        loss = crit(a, b, target=labels.cuda())
        loss.backward()
        print(f"bf16 N=1024 {mode}: loss {loss.item():.6f} |g| {a.grad.norm().item():.3e}")
    # the opt-in backward kernels (128 x 256 S tiles; CTA pair) on the same case
    import ctypes
    from spcl_b200 import _native as nat
    h = nat.lib()
    h.spcl_debug_set_flags.argtypes = [ctypes.c_int]
    for flag, name in ((16384, "bwd_wide_kernel"), (32768, "bwd2_kernel")):
        h.spcl_debug_set_flags(flag)
        try:
            crit = spcl_b200.SelfPacedSupConLoss(weight_update="soft", correct_grad=True, precision="bf16")
            crit.set_gamma(6.0)
            a, b = z1.cuda().requires_grad_(True), z2.cuda().requires_grad_(True)
            crit(a, b, target=labels.cuda()).backward()
            print(f"{name}: |g| {a.grad.norm().item():.3e}")
        finally:
            h.spcl_debug_set_flags(0)
    # ragged N (tail tile) and d = 200 (padded to 256)
    lab = acdc_meta_labels(333)["patient"]
    y1, y2 = make_views(lab, 200, sigma=0.7, seed=2)
    crit = spcl_b200.SupConLoss1(precision="bf16")
    a, b = y1.cuda().requires_grad_(True), y2.cuda().requires_grad_(True)
    crit(a, b, target=lab.tolist()).backward()
    # grouped fp32 launch
    meta = acdc_meta_labels(128)
    crits, feats, tg = [], [], []
    for kind, gm in (("partition", 5.0), ("patient", 3.5)):
        c = spcl_b200.SelfPacedSupConLoss(weight_update="soft", correct_grad=True, precision="fp32")
        c.set_gamma(gm)
        v1, v2 = make_views(meta[kind], 64, sigma=0.7, seed=3)
        crits.append(c); feats.append((v1.cuda().requires_grad_(True), v2.cuda().requires_grad_(True))); tg.append(meta[kind].tolist())
    sum(spcl_b200.grouped_forward(crits, feats, tg)).backward()          # one cooperative launch (fused_group_kernel)
    import os
    os.environ["SPCL_FUSED_SMALL"] = "0"                                  # the staged fp32 route: one launch per stage
    try:
        for (f1, f2) in feats:
            f1.grad = f2.grad = None
        sum(spcl_b200.grouped_forward(crits, feats, tg)).backward()
        c1 = spcl_b200.SelfPacedSupConLoss(weight_update="hard", precision="fp32")
        c1.set_gamma(4.0)
        c1(feats[0][0], feats[0][1], target=tg[0]).backward()
    finally:
        del os.environ["SPCL_FUSED_SMALL"]
    # ragged fused problem: N = 2 x 97 anchors, d = 200 (two column passes of the dZ stage, scalar tail of the vector red)
    labr = acdc_meta_labels(97)["patient"]
    r1, r2 = make_views(labr, 200, sigma=0.7, seed=5)
    cr = spcl_b200.SelfPacedSupConLoss(weight_update="soft", precision="fp32")
    cr.set_gamma(5.0)
    ra, rb = r1.cuda().requires_grad_(True), r2.cuda().requires_grad_(True)
    cr(ra, rb, target=labr.tolist()).backward()
    # l2norm rows / NCHW, fused tail, dense front end
    for shape in ((257, 128), (4, 96, 9, 7), (8, 128, 16, 16)):
        x = torch.randn(*shape, device="cuda", requires_grad=True)
        y, _ = ops.l2norm_fwd(x, 1, 1e-12)
        y.sum().backward()
    c = spcl_b200.SelfPacedSupConLoss(weight_update="soft", precision="bf16")
    c.set_gamma(6.0)
    x1 = torch.randn(8, 128, 8, 8, device="cuda", requires_grad=True)
    x2 = torch.randn(8, 128, 8, 8, device="cuda", requires_grad=True)
    c.forward_raw(x1, x2).backward()
    xd = torch.randn(4, 64, 28, 28, device="cuda", requires_grad=True)
    ops.dense_rows(xd, (7, 7)).sum().backward()
    pts = spcl_b200.point_coordinates(4, 7, 7, 5, seed=0).cuda()
    ops.dense_rows(xd, (7, 7), pts).sum().backward()
    torch.cuda.synchronize()
    print("sanitize_case: done")


if __name__ == "__main__":
    main()
