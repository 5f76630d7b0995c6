"""Dense front end, hook glue (8 f2) and the cfg5 encoder step on the GPU (B200 only).

The front-end kernels are floating point: they are compared with the numpy oracle (oracle/dense_frontend.py, pinned
to the reference-generated fixture tests/golden/dense_points.npz), with that fixture directly, and with plain
torch fp32 (adaptive_avg_pool2d + F.normalize) at sizes the oracle would be slow on.
Tolerances: forward rows 2e-6 absolute (unit vectors; fp32 summation order differs), gradients 2e-5 of max|g|.
"""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import spcl_b200
from spcl_b200 import hooks
from spcl_b200.workloads import acdc_encoder, acdc_meta_labels
from oracle.dense_frontend import dense_rows as oracle_rows, dense_rows_grad as oracle_grad
from oracle.dense_port import dense_supcon
from conftest import Golden

pytestmark = pytest.mark.gpu

DENSE = Golden("dense_points.npz")
ROW_ATOL, GRAD_REL = 2e-6, 2e-5


def _torch_rows(x, ph, pw, points=None):
    y = F.normalize(F.adaptive_avg_pool2d(x, (ph, pw)), p=2, dim=1).permute(0, 2, 3, 1).reshape(x.shape[0], ph * pw, -1)
    if points is not None:
        y = torch.stack([y[b, points[b].long()] for b in range(x.shape[0])])
    return y.reshape(-1, x.shape[1])


@pytest.mark.parametrize("name", DENSE.cases)
def test_dense_rows_match_reference_fixture_and_oracle(name):
    ph, pw, P, seed = (int(v) for v in DENSE[f"{name}/geom"])
    x_np = DENSE[f"{name}/x"]
    pts = torch.from_numpy(DENSE[f"{name}/points"])
    for points, want in ((pts, DENSE[f"{name}/rows"]), (None, DENSE[f"{name}/all_rows"])):
        x = torch.from_numpy(x_np).cuda().requires_grad_(True)
        rows = spcl_b200.ops.dense_rows(x, (ph, pw), points)
        np.testing.assert_allclose(rows.detach().cpu().numpy(), want, rtol=0, atol=ROW_ATOL)
        gy = torch.randn(rows.shape, generator=torch.Generator().manual_seed(1))
        rows.backward(gy.cuda())
        ref = oracle_grad(x_np, ph, pw, gy.numpy().astype(np.float64), None if points is None else points.numpy())
        err = np.abs(x.grad.cpu().numpy() - ref).max() / np.abs(ref).max()
        assert err < GRAD_REL, (name, points is None, err)
    # seeded draw == the reference's FixRandomSeed(seed) draw
    x = torch.from_numpy(x_np).cuda()
    rows = spcl_b200.region_extractor(x, P, spatial_size=(ph, pw), seed=seed)
    np.testing.assert_allclose(rows.cpu().numpy(), DENSE[f"{name}/rows"], rtol=0, atol=ROW_ATOL)


@pytest.mark.parametrize("B,C,H,W,ph,pw", [
    (2, 128, 32, 32, 32, 32),        # cfg3 geometry: no pooling, pure normalise + NCHW -> rows
    (3, 64, 28, 28, 16, 16),         # overlapping windows (28 / 16)
    (2, 128, 56, 56, 10, 10),        # decoder default spatial_size
    (1, 256, 224, 224, 32, 32),      # widest embedding, full-resolution map, 7 x 7 windows
    (2, 33, 17, 19, 5, 7),           # ragged everything
    (1, 8, 5, 5, 1, 1),              # global pooling
])
def test_dense_rows_all_pixels_vs_torch(B, C, H, W, ph, pw):
    gen = torch.Generator().manual_seed(B * 1000 + C)
    x0 = torch.randn(B, C, H, W, generator=gen).cuda()
    x = x0.clone().requires_grad_(True)
    xr = x0.clone().requires_grad_(True)
    rows = spcl_b200.ops.dense_rows(x, (ph, pw))
    ref = _torch_rows(xr, ph, pw)
    assert rows.shape == ref.shape == (B * ph * pw, C)
    assert (rows - ref).abs().max().item() < ROW_ATOL
    gy = torch.randn(rows.shape, generator=gen).cuda()
    rows.backward(gy)
    ref.backward(gy)
    err = (x.grad - xr.grad).abs().max().item() / xr.grad.abs().max().item()
    assert err < GRAD_REL, err


@pytest.mark.parametrize("B,C,H,W,ph,pw,P", [(8, 128, 56, 56, 10, 10, 5), (4, 256, 28, 28, 28, 28, 9),
                                            (3, 40, 23, 31, 6, 8, 6)])
def test_dense_rows_points_vs_torch(B, C, H, W, ph, pw, P):
    gen = torch.Generator().manual_seed(P)
    x0 = torch.randn(B, C, H, W, generator=gen).cuda()
    pts = spcl_b200.point_coordinates(B, ph, pw, P, seed=11)
    x = x0.clone().requires_grad_(True)
    xr = x0.clone().requires_grad_(True)
    rows = spcl_b200.ops.dense_rows(x, (ph, pw), pts)
    ref = _torch_rows(xr, ph, pw, pts.cuda())
    assert (rows - ref).abs().max().item() < ROW_ATOL
    gy = torch.randn(rows.shape, generator=gen).cuda()
    rows.backward(gy)
    ref.backward(gy)
    assert (x.grad - xr.grad).abs().max().item() / xr.grad.abs().max().item() < GRAD_REL
    with pytest.raises(IndexError):
        spcl_b200.ops.dense_rows(x0, (ph, pw), torch.full((B, P), ph * pw, dtype=torch.int32))


@pytest.mark.parametrize("B,C,H,W,ph,pw,P", [(2, 128, 56, 56, 10, 10, None), (3, 64, 28, 28, 16, 16, None),
                                            (2, 33, 17, 19, 5, 7, None), (4, 96, 56, 56, 10, 10, 5),
                                            (1, 8, 5, 5, 1, 1, None)])
def test_dense_rows_adaptive_max_vs_torch(B, C, H, W, ph, pw, P):
    """pool_name="adaptive_max" (nn.py:57-58): rows and gradients vs torch's adaptive_max_pool2d + F.normalize."""
    gen = torch.Generator().manual_seed(B * 100 + C)
    x0 = torch.randn(B, C, H, W, generator=gen).cuda()
    pts = None if P is None else spcl_b200.point_coordinates(B, ph, pw, P, seed=3)
    x = x0.clone().requires_grad_(True)
    xr = x0.clone().requires_grad_(True)
    rows = spcl_b200.ops.dense_rows(x, (ph, pw), pts, pool="max")
    pooled = F.normalize(F.adaptive_max_pool2d(xr, (ph, pw)), dim=1)
    ref = pooled.permute(0, 2, 3, 1).reshape(B, ph * pw, C)
    if pts is not None:
        ref = torch.gather(ref, 1, pts.cuda().long()[:, :, None].expand(-1, -1, C))
    ref = ref.reshape(-1, C)
    assert rows.shape == ref.shape
    assert (rows - ref).abs().max().item() < ROW_ATOL
    gy = torch.randn(rows.shape, generator=gen).cuda()
    rows.backward(gy)
    ref.backward(gy)
    assert (x.grad - xr.grad).abs().max().item() / xr.grad.abs().max().item() < GRAD_REL
    tail = spcl_b200.DenseProjectionTail((ph, pw), pool_name="adaptive_max")
    assert torch.equal(tail(x0, points=pts), rows.detach())


def test_dense_rows_zero_vector_uses_eps_like_f_normalize():
    x = torch.zeros(1, 16, 8, 8, device="cuda")
    x[0, :, 4:, :] = 1.0
    rows = spcl_b200.ops.dense_rows(x, (2, 2))
    assert torch.equal(rows[:2], torch.zeros_like(rows[:2]))            # 0 / max(0, eps) = 0, no NaN
    assert torch.allclose(rows[2:].norm(dim=1), torch.ones(2, device="cuda"))


# ---- hooks ---------------------------------------------------------------------------------------------------
def _reference_heads(head):
    """the same weights run through plain torch ops (what the reference's heads.py composes)."""
    return copy.deepcopy(head)


def test_encoder_hook_equals_reference_pipeline():
    torch.manual_seed(0)
    n, cin = 24, 64
    labels = acdc_meta_labels(n)
    part = [str(int(v)) for v in labels["partition"]]
    group = [f"patient{int(p):03d}_{'00' if int(c) == 0 else '01'}" for p, c in zip(labels["patient"], labels["cycle"])]
    head = hooks.ProjectionHead(input_dim=cin, hidden_dim=256, output_dim=256, head_type="mlp", normalize=True).cuda()
    ref_head = _reference_heads(head)
    sched = hooks.SelfPacedGammaSchedule(mode="soft", begin_value=8.0, end_value=20.0, max_epoch=3, correct_grad=True,
                                         precision="fp32")
    sched.new_epoch()
    seen = []
    hook = hooks.SPINFONCEEpochHook(name="sp", weight=0.5, projector=head, criterion=sched.criterion,
                                    label_generator=lambda **kw: hooks.get_label("patient", "acdc", **kw),
                                    figure_fn=lambda t, tag: seen.append((tag, tuple(t.shape))))
    fa = torch.randn(n, cin, 7, 7, device="cuda")
    fb = fa + 0.3 * torch.randn_like(fa)
    total = 0.0
    for step in range(2):
        loss = hook(fa, fb, partition_group=part, label_group=group)
        loss.backward()
        # reference pipeline (infonce.py:171-195) on the same weights with torch ops and the dense port
        seq = ref_head._header
        z = seq[:-1](torch.cat([fa, fb]))
        za, zb = torch.chunk(F.normalize(z, p=2, dim=1), 2)
        ref = dense_supcon(za, zb, target=hooks.get_label("patient", "acdc", part, group), gamma=8.0, mode="soft",
                           correct_grad=True)
        (ref.loss * 0.5).backward()
        assert loss.item() == pytest.approx(ref.loss.item() * 0.5, rel=2e-5)
        total += ref.loss.item()
    for p, q in zip(head.parameters(), ref_head.parameters()):
        assert (p.grad - q.grad).abs().max().item() <= 1e-4 * q.grad.abs().max().item() + 1e-9
    s = hook.summary()
    assert s["loss"] == pytest.approx(total / 2, rel=2e-5)
    assert s["age_param"] == pytest.approx(8.0)
    assert 0.0 < s["sp_weight"] <= 1.0 and s["sp_weight"] == pytest.approx(ref.ratio, rel=1e-4)
    assert [t for t, _ in seen] == ["pos_mask", "sim_exp", "sim_logits", "sp_mask"]
    assert all(shape == (2 * n, 2 * n) for _, shape in seen)


def test_dense_hook_equals_reference_pipeline():
    torch.manual_seed(1)
    # fp32 on both sides: the head runs its last 1x1 convolution on the pooled rows (F.linear), the reference order runs
    # it at full resolution through cuDNN, which would otherwise pick TF32 for one arm only
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    b, cin = 6, 32
    head = hooks.DenseProjectionHead(input_dim=cin, hidden_dim=64, output_dim=128, head_type="mlp", normalize=True,
                                     spatial_size=(10, 10)).cuda()
    ref_head = copy.deepcopy(head)
    hook = hooks.INFONCEDenseHook(name="dense", weight=1.0, projector=head,
                                  criterion=spcl_b200.SupConLoss1(precision="fp32"), point_nums=5)
    fa = torch.randn(b, cin, 28, 28, device="cuda")
    fb = fa + 0.2 * torch.randn_like(fa)
    loss = hook(fa, fb, seed=3)
    loss.backward()
    out = F.normalize(F.adaptive_avg_pool2d(ref_head._projector(torch.cat([fa, fb])), (10, 10)), p=2, dim=1)
    pts = spcl_b200.point_coordinates(b, 10, 10, 5, seed=3).long().cuda()
    flat = out.flatten(2)                                              # [2b, C, 100]
    sel = torch.stack([flat[i][:, pts[i % b]].t() for i in range(2 * b)]).reshape(2 * b * 5, -1)
    za, zb = torch.chunk(sel, 2)
    ref = dense_supcon(za, zb, target=list(range(b * 5)), mode="none")
    ref.loss.backward()
    assert loss.item() == pytest.approx(ref.loss.item(), rel=2e-5)
    for p, q in zip(head.parameters(), ref_head.parameters()):
        assert (p.grad - q.grad).abs().max().item() <= 1e-4 * q.grad.abs().max().item() + 1e-9
    # forward() keeps the reference's [B, C, ph, pw] output
    assert torch.allclose(head(fa), out[:b], atol=ROW_ATOL)
    # the reference's order of operations (convolution at full resolution, then pool) is still available and agrees
    head.commute_pooling = False
    assert torch.allclose(head(fa), out[:b], atol=ROW_ATOL)


# ---- cfg5: encoder pre-training step, fused loss drop-in vs the dense port on the same weights ---------------------
def test_cfg5_encoder_step_fused_vs_reference_loss():
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    n = 16                                                   # test-sized batch (bench harness: tools/cfg5_step.py)
    labels = acdc_meta_labels(n)["partition"].tolist()

    def make():
        torch.manual_seed(0)
        enc = acdc_encoder(1, 256).cuda()
        head = hooks.ProjectionHead(input_dim=256, hidden_dim=256, output_dim=256, head_type="mlp", normalize=True).cuda()
        # SGD, not Adam: Adam's sign-like normalisation would amplify rounding-level gradient differences
        opt = torch.optim.SGD(list(enc.parameters()) + list(head.parameters()), lr=1e-2, momentum=0.9)
        return enc, head, opt

    arms = {"fused": make(), "ref": make()}
    crit = spcl_b200.SelfPacedSupConLoss(weight_update="soft", correct_grad=True)
    crit.set_gamma(8.0)
    gen = torch.Generator().manual_seed(5)
    for step in range(3):
        x = torch.randn(2 * n, 1, 64, 64, generator=gen).cuda()
        vals = {}
        for name, (enc, head, opt) in arms.items():
            opt.zero_grad(set_to_none=True)
            if name == "fused":
                za, zb = torch.chunk(head(enc(x)), 2)
                loss = crit(za, zb, target=labels)
            else:
                z = head._header[:-1](enc(x))
                za, zb = torch.chunk(F.normalize(z, p=2, dim=1), 2)
                loss = dense_supcon(za, zb, target=labels, gamma=8.0, mode="soft", correct_grad=True).loss
            loss.backward()
            gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in list(enc.parameters()) + list(head.parameters())))
            opt.step()
            vals[name] = (loss.item(), gn.item())
        # identical weights and inputs: the two arms may only drift by fp32 rounding through the optimiser
        assert vals["fused"][0] == pytest.approx(vals["ref"][0], rel=1e-3), (step, vals)
        assert vals["fused"][1] == pytest.approx(vals["ref"][1], rel=2e-2), (step, vals)


def _ref_parts():
    from baseline import ref_loader
    if not ref_loader.available():
        pytest.skip("baseline/_ref is not installed (tools/install_ref.sh)")
    return ref_loader.unet_module().UNet, ref_loader.heads_module(), ref_loader.loss_module()


def test_cfg5_step_on_the_reference_unet_matches_the_reference_loss():
    """cfg5 with the reference's OWN modules on both sides of the loss: UNet(..., until="Conv5") and
    ProjectionHead(256, 256, 256, "mlp") from baseline/_ref; the only difference between the arms is the loss module
    (unmodified contrast_loss3.SelfPacedSupConLoss vs the fused one)."""
    UNet, heads, ref_loss = _ref_parts()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    n = 12
    labels = acdc_meta_labels(n)["partition"].tolist()

    def make():
        torch.manual_seed(0)
        enc = UNet(input_dim=1, num_classes=4, max_channel=256).cuda()
        head = heads.ProjectionHead(input_dim=256, hidden_dim=256, output_dim=256, head_type="mlp", normalize=True).cuda()
        params = [q for q in list(enc.parameters()) + list(head.parameters())]
        return enc, head, torch.optim.SGD(params, lr=1e-2, momentum=0.9)

    arms = {"fused": make(), "ref": make()}
    crits = {"fused": spcl_b200.SelfPacedSupConLoss(weight_update="soft", correct_grad=True),
             "ref": ref_loss.SelfPacedSupConLoss(weight_update="soft", correct_grad=True)}
    for c in crits.values():
        c.set_gamma(8.0)
    gen = torch.Generator().manual_seed(5)
    for step in range(3):
        x = torch.randn(2 * n, 1, 96, 96, generator=gen).cuda()
        vals = {}
        for name, (enc, head, opt) in arms.items():
            opt.zero_grad(set_to_none=True)
            za, zb = torch.chunk(head(enc(x, until="Conv5")), 2)
            loss = crits[name](za, zb, target=labels)
            loss.backward()
            gn = torch.sqrt(sum((q.grad.double() ** 2).sum() for q in head.parameters()))
            opt.step()
            vals[name] = (loss.item(), gn.item(), crits[name].downgrade_ratio)
        assert vals["fused"][0] == pytest.approx(vals["ref"][0], rel=1e-3), (step, vals)
        assert vals["fused"][1] == pytest.approx(vals["ref"][1], rel=2e-2), (step, vals)
        assert vals["fused"][2] == pytest.approx(vals["ref"][2], rel=1e-3), (step, vals)


@pytest.mark.parametrize("pool_name", ["adaptive_avg", "adaptive_max"])
def test_decoder_step_on_the_reference_unet(pool_name):
    """Decoder-stage step (main_pretrain_decoder.py:42-76 + infonce.py:198-241): reference UNet up to Up_conv3 with the
    encoder frozen, the reference's DenseProjectionHead + region_extractor + SupConLoss1 on one side, this repo's head
    tail (pool + normalise + point gather in one kernel) + fused loss on the other; same weights, same sampled points."""
    import importlib.util
    import pathlib
    spec = importlib.util.spec_from_file_location("decoder_step", pathlib.Path(__file__).resolve().parent.parent / "tools" / "decoder_step.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = mod.measure(batch=6, steps=3, warmup=0, image=96, pool_name=pool_name, tf32=False)
    if "reference" not in out:
        pytest.skip("baseline/_ref is not installed (tools/install_ref.sh)")
    for lf, lr in zip(out["fused"]["losses"], out["reference"]["losses"]):
        assert lf == pytest.approx(lr, rel=1e-4), out
    for gf, gr in zip(out["fused"]["grad_norms"], out["reference"]["grad_norms"]):
        assert gf == pytest.approx(gr, rel=5e-3), out


# ---- grouped launch: K problems per launch == K separate calls -----------------------------
@pytest.mark.parametrize("shapes", [
    [(256, 256, "partition"), (256, 256, "patient"), (256, 256, "cycle")],            # cfg2
    [(24, 64, "patient"), (100, 128, "composite"), (256, 256, "cycle"), (7, 16, "self")],   # ragged group
    [(64, 128, "partition")],
])
@pytest.mark.parametrize("graph", [False, True])
def test_grouped_launch_equals_separate_calls(shapes, graph):
    from spcl_b200.workloads import make_views
    specs = [("soft", 5.0, True), ("hard", 3.5, False), ("soft", 2.0, False), ("none", 1e6, False)]
    crits_a, crits_b, feats_a, feats_b, targets = [], [], [], [], []
    for k, (n, d, kind) in enumerate(shapes):
        lab = acdc_meta_labels(n)[kind]
        z1, z2 = make_views(lab, d, sigma=0.7, seed=k)
        mode, gamma, cg = specs[k % len(specs)]
        for crits, feats in ((crits_a, feats_a), (crits_b, feats_b)):
            if mode == "none":
                c = spcl_b200.SupConLoss1(precision="fp32")
            else:
                c = spcl_b200.SelfPacedSupConLoss(weight_update=mode, correct_grad=cg, precision="fp32")
                c.set_gamma(gamma)
            crits.append(c)
            feats.append((z1.cuda().requires_grad_(True), z2.cuda().requires_grad_(True)))
        targets.append(lab.tolist() if k % 2 else lab.int().cuda())
    weights = [1.0, 0.5, 0.25, 2.0][:len(shapes)]
    for rep in range(2 if graph else 1):                  # the second pass replays the cached graph
        for a, b in feats_a:
            a.grad = b.grad = None
        grouped = spcl_b200.grouped_forward(crits_a, feats_a, targets, cuda_graph=graph)
        sum(w * l for w, l in zip(weights, grouped)).backward()
    single = [c(z1, z2, target=t) for c, (z1, z2), t in zip(crits_b, feats_b, targets)]
    sum(w * l for w, l in zip(weights, single)).backward()
    for k in range(len(shapes)):
        assert grouped[k].item() == pytest.approx(single[k].item(), rel=1e-5), k
        if hasattr(crits_a[k], "downgrade_ratio"):
            assert crits_a[k].downgrade_ratio == pytest.approx(crits_b[k].downgrade_ratio, rel=1e-5)
        for ga, gb in zip(feats_a[k], feats_b[k]):
            scale = gb.grad.abs().max().item()
            assert (ga.grad - gb.grad).abs().max().item() <= 2e-5 * scale + 1e-12, k


def test_grouped_launch_falls_back_to_separate_calls_for_tensor_core_sizes():
    from spcl_b200.workloads import make_views
    lab = torch.arange(1024) // 8
    z1, z2 = make_views(lab, 128, sigma=0.7, seed=0)
    c1, c2 = spcl_b200.SupConLoss1(), spcl_b200.SupConLoss1()
    a = spcl_b200.grouped_forward([c1], [(z1.cuda(), z2.cuda())], [lab.tolist()])[0]          # N = 2048: bf16 path
    b = c2(z1.cuda(), z2.cuda(), target=lab.tolist())
    assert a.item() == pytest.approx(b.item(), rel=1e-6)


def test_subsampled_figures_match_the_full_diagnostics():
    from spcl_b200.workloads import make_views
    n = 96
    lab = acdc_meta_labels(n)["patient"]
    z1, z2 = make_views(lab, 64, sigma=0.7, seed=2)
    crit = spcl_b200.SelfPacedSupConLoss(weight_update="soft", precision="fp32")
    crit.set_gamma(6.0)
    crit(z1.cuda(), z2.cuda(), target=lab.tolist())
    idx = torch.linspace(0, 2 * n - 1, 50).round().long().unique().cuda()
    for name in ("sim_exp", "sim_logits", "pos_mask", "neg_mask", "sp_mask"):
        full = getattr(crit, name)
        assert torch.equal(crit.figure(name, max_side=4096), full)               # small N: the full matrix
        sub = crit.figure(name, max_side=50)
        assert sub.shape == (idx.numel(), idx.numel())
        assert torch.allclose(sub, full[idx][:, idx], atol=1e-5), name
    # beyond the full-matrix limit only the figure form is available
    big = torch.nn.functional.normalize(torch.randn(9000, 32, device="cuda"), dim=1)
    c2 = spcl_b200.SupConLoss1()
    c2(big, big.clone())
    with pytest.raises(RuntimeError):
        c2.sim_exp
    assert c2.figure("sim_exp", 256).shape[0] <= 256
