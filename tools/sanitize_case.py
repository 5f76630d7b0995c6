"""Workload of tools/sanitize.sh: one small pass over every kernel family under compute-sanitizer.

smoke() (fp32 SIMT + bf16 tensor-core fwd/bwd, n = 192), a 1024-anchor bf16 case with slice labels in hard and soft
mode (tcgen05 / TMA / mbarrier pipelines of stats_kernel, sp_kernel, bwd_kernel, transpose_kernel), the grouped fp32
launch, l2norm rows / NCHW fwd + bwd, the fused projector tail and the dense front end (pool rows / points)."""
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
    sum(spcl_b200.grouped_forward(crits, feats, tg)).backward()
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
