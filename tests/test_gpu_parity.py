"""Parity of the CUDA paths against the oracle / the reference's golden vectors (B200 only)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import spcl_b200
from spcl_b200 import _native as nat
from spcl_b200.workloads import acdc_meta_labels, make_views, make_workload
from oracle.closed_form import supcon_closed_form
from oracle.dense_port import dense_supcon
from conftest import Golden, excl_case_inputs, parse_cfg1_case
from torch_ref import supcon_ref64

pytestmark = pytest.mark.gpu

CFG1 = Golden("cfg1_n64_d128.npz")
TINY = Golden("tiny_n5_d16.npz")
CFG2 = Golden("cfg2_n256_d256.npz")
EXCL = Golden("excl_cases.npz")
MODE = {"none": nat.MODE_NONE, "hard": nat.MODE_HARD, "soft": nat.MODE_SOFT}

# ---- stated tolerances -------------------------------------------------------------------------
# fp32 SIMT path vs the fp32 reference: both round in fp32, only the summation order differs.
FP32_LOSS_RTOL, FP32_GRAD_REL = 2e-5, 1e-4
# bf16 tensor-core path vs the fp64 oracle evaluated on the SAME bf16-rounded operands: remaining error is
# ex2.approx (2^-22), fp32 accumulation order and the bf16 rounding of T before the second MMA (2^-9 per term).
BF16_TIGHT_LOSS_RTOL, BF16_TIGHT_GRAD_REL, BF16_TIGHT_COS = 3e-4, 2e-2, 0.9999
# bf16 path vs the fp32 reference on raw fp32 inputs: adds the operand quantisation (|ds| <= 2^-8/tau).
BF16_LOOSE_LOSS_RTOL, BF16_LOOSE_COS = 2e-2, 0.999


def _run(z1, z2, *, cls="SP", target=None, mask=None, gamma=1e6, mode="hard", correct_grad=False,
         temperature=0.07, precision="fp32", validate=True):
    a = torch.as_tensor(z1).cuda().requires_grad_(True)
    b = torch.as_tensor(z2).cuda().requires_grad_(True)
    if cls in ("SupConLoss1", "SupConLoss1Excl"):
        crit = spcl_b200.SupConLoss1(temperature=temperature, exclude_other_pos=cls == "SupConLoss1Excl",
                                     precision=precision, validate=validate)
    else:
        crit = spcl_b200.SelfPacedSupConLoss(temperature=temperature, weight_update=mode, correct_grad=correct_grad,
                                             precision=precision, validate=validate)
        crit.set_gamma(gamma)
    kw = {}
    if mask is not None:
        kw["mask"] = torch.as_tensor(mask).cuda()
    elif target is not None:
        kw["target"] = torch.as_tensor(target).cuda() if isinstance(target, np.ndarray) else target
    loss = crit(a, b, **kw)
    loss.backward()
    ratio = crit.downgrade_ratio if cls == "SP" else float("nan")
    return dict(loss=loss.item(), ratio=ratio, dz1=a.grad.cpu().numpy(), dz2=b.grad.cpu().numpy(), crit=crit)


def _grad_metrics(res, ref):
    g = np.concatenate([res["dz1"], res["dz2"]]).astype(np.float64)
    r = np.concatenate([np.asarray(ref["dz1"]), np.asarray(ref["dz2"])]).astype(np.float64)
    if np.abs(r).max() == 0.0:           # e.g. hard weighting with every positive above gamma: zero gradient
        return (0.0, 1.0) if np.abs(g).max() < 1e-12 else (np.inf, 0.0)
    rel = np.abs(g - r).max() / np.abs(r).max()
    cos = (g * r).sum() / (np.linalg.norm(g) * np.linalg.norm(r))
    return rel, cos


# ------------------------------------------------------------------------------------------------
# fp32 path vs the reference's own outputs
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", EXCL.cases)
def test_exclude_other_pos_matches_reference(name):
    """SupConLoss1(exclude_other_pos=True), contrast_loss3.py:97-100, against the reference's own outputs."""
    z1, z2, kw = excl_case_inputs(EXCL, name)
    res = _run(z1, z2, cls="SupConLoss1Excl", precision="auto", **kw)
    ref = EXCL.case(name)
    assert np.isclose(res["loss"], ref["loss"], rtol=FP32_LOSS_RTOL), (res["loss"], ref["loss"])
    rel, _ = _grad_metrics(res, ref)
    assert rel <= FP32_GRAD_REL, rel


def test_exclude_other_pos_large_n_uses_fp32_kernels():
    """N = 2048 >= the auto threshold still runs (fp32 kernels) and matches the fp64 oracle; bf16 is refused."""
    labels = acdc_meta_labels(1024)["patient"]
    z1, z2 = make_views(labels, 64, sigma=0.7, seed=4)
    res = _run(z1, z2, cls="SupConLoss1Excl", target=labels.tolist(), precision="auto")
    ref = supcon_closed_form(z1.numpy(), z2.numpy(), target=labels.tolist(), mode="excl")
    assert np.isclose(res["loss"], ref["loss"], rtol=FP32_LOSS_RTOL)
    rel, _ = _grad_metrics(res, ref)
    assert rel <= FP32_GRAD_REL, rel
    with pytest.raises(nat.SpclError):
        _run(z1, z2, cls="SupConLoss1Excl", target=labels.tolist(), precision="bf16")


@pytest.mark.parametrize("name", CFG1.cases)
def test_fp32_path_matches_reference_cfg1(name):
    kw = parse_cfg1_case(name, CFG1)
    res = _run(CFG1["z1"], CFG1["z2"], precision="fp32", **kw)
    ref = CFG1.case(name)
    assert np.isclose(res["loss"], ref["loss"], rtol=FP32_LOSS_RTOL), (res["loss"], ref["loss"])
    if kw["cls"] != "SupConLoss1":
        assert np.isclose(res["ratio"], ref["ratio"], rtol=1e-5, atol=1e-7)
    rel, _ = _grad_metrics(res, ref)
    assert rel < FP32_GRAD_REL, rel


@pytest.mark.parametrize("name", CFG2.cases)
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_cfg2_matches_reference(name, precision):
    z1, z2 = CFG2[f"{name}/z1"], CFG2[f"{name}/z2"]
    res = _run(z1, z2, target=CFG2[f"{name}/labels"].tolist(), gamma=float(CFG2[f"{name}/gamma"]), mode="soft",
               precision=precision)
    ref = CFG2.case(name)
    rel, cos = _grad_metrics(res, ref)
    if precision == "fp32":
        assert np.isclose(res["loss"], ref["loss"], rtol=FP32_LOSS_RTOL)
        assert rel < FP32_GRAD_REL
    else:
        assert np.isclose(res["loss"], ref["loss"], rtol=BF16_LOOSE_LOSS_RTOL)
        assert cos > BF16_LOOSE_COS, cos


@pytest.mark.parametrize("name", TINY.cases)
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_tiny_ragged_batch(name, precision):
    labels = TINY["labels"].tolist()
    kw = {"sp_soft_g3": dict(target=labels, gamma=3.0, mode="soft"),
          "sp_hard_g3": dict(target=labels, gamma=3.0, mode="hard"),
          "supcon1": dict(cls="SupConLoss1", target=labels),
          "sp_simclr": dict(gamma=4.0, mode="soft")}[name]
    res = _run(TINY["z1"], TINY["z2"], precision=precision, **kw)
    ref = TINY.case(name)
    if precision == "fp32":
        assert np.isclose(res["loss"], ref["loss"], rtol=FP32_LOSS_RTOL)
        assert _grad_metrics(res, ref)[0] < FP32_GRAD_REL
    else:
        assert np.isclose(res["loss"], ref["loss"], rtol=5e-2, atol=5e-3)
        assert _grad_metrics(res, ref)[1] > 0.99


# ------------------------------------------------------------------------------------------------
# bf16 tensor-core path
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", [c for c in CFG1.cases if "trimask" not in c])
def test_bf16_path_cfg1(name):
    kw = parse_cfg1_case(name, CFG1)
    res = _run(CFG1["z1"], CFG1["z2"], precision="bf16", **kw)
    # (a) tight: oracle on the bf16-rounded operands
    zb1 = torch.from_numpy(CFG1["z1"]).bfloat16().float().numpy()
    zb2 = torch.from_numpy(CFG1["z2"]).bfloat16().float().numpy()
    okw = {k: v for k, v in kw.items() if k != "cls"}
    tight = supcon_closed_form(zb1, zb2, **okw)
    ref = CFG1.case(name)
    if kw["mode"] == "hard" and kw["gamma"] < 1e5:
        # a pair whose loss sits within rounding of gamma may flip: allow two flips, state it
        n_pos = tight["c"].sum()
        assert abs(res["ratio"] - tight["ratio"]) <= 2.0 / n_pos + 1e-6
        assert np.isclose(res["loss"], tight["loss"], rtol=2e-2)
    else:
        assert np.isclose(res["loss"], tight["loss"], rtol=BF16_TIGHT_LOSS_RTOL), (res["loss"], tight["loss"])
        if kw["cls"] != "SupConLoss1":
            assert np.isclose(res["ratio"], tight["ratio"], rtol=1e-4, atol=1e-6)
        rel, cos = _grad_metrics(res, tight)
        assert rel < BF16_TIGHT_GRAD_REL and cos > BF16_TIGHT_COS, (rel, cos)
    # (b) loose: the reference on the raw fp32 inputs
    assert np.isclose(res["loss"], ref["loss"], rtol=BF16_LOOSE_LOSS_RTOL)
    assert _grad_metrics(res, ref)[1] > BF16_LOOSE_COS


@pytest.mark.parametrize("n,d,kind", [(75, 96, "partition"), (192, 64, "patient"), (129, 200, "cycle"),
                                      (640, 128, "composite"), (300, 256, "self")])
@pytest.mark.parametrize("mode", ["none", "soft"])
def test_bf16_path_shapes(n, d, kind, mode):
    labels = acdc_meta_labels(n)[kind]
    z1, z2 = make_views(labels, d, sigma=0.7, seed=n)
    z1, z2 = z1.bfloat16().float(), z2.bfloat16().float()
    cls = "SupConLoss1" if mode == "none" else "SP"
    res = _run(z1, z2, cls=cls, target=labels.tolist(), gamma=6.0, mode=mode, correct_grad=(mode == "soft"),
               precision="bf16", validate=False)
    ref = supcon_closed_form(z1.numpy(), z2.numpy(), target=labels.tolist(), gamma=6.0, mode=mode,
                             correct_grad=(mode == "soft"))
    assert np.isclose(res["loss"], ref["loss"], rtol=BF16_TIGHT_LOSS_RTOL), (res["loss"], ref["loss"])
    rel, cos = _grad_metrics(res, ref)
    assert rel < BF16_TIGHT_GRAD_REL and cos > BF16_TIGHT_COS, (rel, cos)


def test_bf16_and_fp32_paths_agree_n4096():
    z1, z2, labels = make_workload("cfg3_dense_2x16384_d128_slice")
    z1, z2, labels = z1[:2048].bfloat16().float(), z2[:2048].bfloat16().float(), labels[:2048] // 8
    a = _run(z1, z2, target=labels.int().numpy(), gamma=8.0, mode="soft", precision="fp32", validate=False)
    b = _run(z1, z2, target=labels.int().numpy(), gamma=8.0, mode="soft", precision="bf16", validate=False)
    assert np.isclose(a["loss"], b["loss"], rtol=BF16_TIGHT_LOSS_RTOL)
    assert np.isclose(a["ratio"], b["ratio"], rtol=1e-4)
    rel, cos = _grad_metrics(b, a)
    assert rel < BF16_TIGHT_GRAD_REL and cos > BF16_TIGHT_COS


# ------------------------------------------------------------------------------------------------
# BASELINE full size (cfg3: N = 32768, d = 128): fp64 torch reference on the GPU + size-independent properties
# ------------------------------------------------------------------------------------------------
def test_torch_ref_matches_oracle():
    kw = dict(gamma=5.0, mode="soft", correct_grad=True)
    lab = torch.from_numpy(CFG1["labels_patient"])
    ref = supcon_closed_form(CFG1["z1"], CFG1["z2"], target=lab.tolist(), **kw)
    out = supcon_ref64(torch.from_numpy(CFG1["z1"]).cuda(), torch.from_numpy(CFG1["z2"]).cuda(), lab.cuda(), **kw)
    assert np.isclose(out["loss"], ref["loss"], rtol=1e-10)
    np.testing.assert_allclose(out["dz1"].cpu().numpy(), ref["dz1"], rtol=1e-8, atol=1e-14)


@pytest.mark.parametrize("workload,mode,gamma", [
    ("cfg3_dense_2x16384_d128_simclr", "soft", 8.0),
    ("cfg3_dense_2x16384_d128_slice", "soft", 10.0),
    ("cfg3_dense_2x16384_d128_slice", "none", 1e6),
    ("cfg3_dense_2x16384_d128_slice", "hard", 10.0),      # a real threshold: ~all positives kept, near-gamma pairs exist
    ("cfg3_dense_2x16384_d128_slice", "hard", 8.5),       # threshold inside the mass of l_ij (ratio well below 1)
])
def test_cfg3_full_size_against_fp64(workload, mode, gamma):
    z1, z2, labels = make_workload(workload)
    z1, z2 = z1.bfloat16().float().cuda(), z2.bfloat16().float().cuda()
    lab = labels.int().cuda()
    cls = "SupConLoss1" if mode == "none" else "SP"
    res = _run(z1, z2, cls=cls, target=lab, gamma=gamma, mode=mode, precision="bf16", validate=False)
    ref = supcon_ref64(z1, z2, lab, gamma=gamma, mode=mode)
    # hard mode: a pair whose l_ij is within fp32 / ex2.approx rounding of gamma may land on the other side of the
    # threshold than in fp64; each flip moves the loss by l_ij / (N c_i) ~ 1e-8 relative and one count of ~3e7
    # positives, so the same tolerances hold -- the gradient bound below is the one that would see a real error
    assert np.isclose(res["loss"], ref["loss"], rtol=BF16_TIGHT_LOSS_RTOL), (res["loss"], ref["loss"])
    if mode != "none":
        assert np.isclose(res["ratio"], ref["ratio"], rtol=2e-4), (res["ratio"], ref["ratio"])
    logD = res["crit"]._diag.row_stats[0, : 2 * z1.shape[0]].double()
    assert (logD - ref["logD"]).abs().max().item() < 2e-4
    ref_np = dict(dz1=ref["dz1"].cpu().numpy(), dz2=ref["dz2"].cpu().numpy())
    rel, cos = _grad_metrics(res, ref_np)
    assert rel < 3e-2 and cos > BF16_TIGHT_COS, (rel, cos)


def test_cfg3_properties():
    z1, z2, labels = make_workload("cfg3_dense_2x16384_d128_simclr")
    n, d = z1.shape
    # (1) gamma -> inf (hard) == SupConLoss1   (reference __main__ identity, contrast_loss2.py:330-346)
    a = _run(z1, z2, target=labels.int().numpy(), gamma=1e6, mode="hard", precision="bf16")
    b = _run(z1, z2, cls="SupConLoss1", target=labels.int().numpy(), precision="bf16")
    assert np.isclose(a["loss"], b["loss"], rtol=1e-5) and a["ratio"] == 1.0
    assert _grad_metrics(a, b)[0] < 1e-3
    # (2) target=range(n) == no target (SimCLR)
    c = _run(z1, z2, gamma=1e6, mode="hard", precision="bf16")
    assert np.isclose(a["loss"], c["loss"], rtol=1e-6)
    # (3) rotation invariance of f(Z Z^T)  =>  Z^T dZ is symmetric
    Z = np.concatenate([z1.numpy(), z2.numpy()]).astype(np.float64)
    G = np.concatenate([a["dz1"], a["dz2"]]).astype(np.float64)
    M = Z.T @ G
    assert np.abs(M - M.T).max() <= 2e-2 * np.abs(M).max()
    # (4) anchor permutation equivariance
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(5))
    p = _run(z1[perm], z2[perm], target=labels[perm].int().numpy(), gamma=1e6, mode="hard", precision="bf16")
    assert np.isclose(a["loss"], p["loss"], rtol=1e-5)
    assert np.abs(p["dz1"] - a["dz1"][perm.numpy()]).max() <= 2e-2 * np.abs(a["dz1"]).max()


# ------------------------------------------------------------------------------------------------
# edge cases and error behaviour
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_single_pair_batch(precision):
    z = F.normalize(torch.randn(1, 32, generator=torch.Generator().manual_seed(1)), dim=1)
    res = _run(z, z.clone(), precision=precision, gamma=3.0, mode="soft")
    assert abs(res["loss"]) < 1e-5 and np.isfinite(res["dz1"]).all()


def test_zero_positive_row_raises_runtime_error():
    n = 8
    z1, z2 = make_views(torch.arange(n) % 2, 32, seed=3)
    tri = torch.ones(n, n)
    tri[0, :] = 0.0                      # anchor 0 (and n) has no positive at all -> 0/0 (:196, :203)
    crit = spcl_b200.SupConLoss1()
    with pytest.raises(RuntimeError):
        crit(z1.cuda(), z2.cuda(), mask=tri.cuda())


def test_trimask_needs_fp32_path():
    z1, z2 = make_views(torch.arange(8) % 2, 32, seed=3)
    crit = spcl_b200.SupConLoss1(precision="bf16")
    with pytest.raises(nat.SpclError):
        crit(z1.cuda(), z2.cuda(), mask=torch.ones(8, 8).cuda())


def test_unnormalised_input_asserts():
    crit = spcl_b200.SupConLoss1()
    with pytest.raises(AssertionError):
        crit(torch.randn(8, 16).cuda(), torch.randn(8, 16).cuda())


def test_non_contiguous_and_half_inputs():
    labels = acdc_meta_labels(64)["partition"]
    z1, z2 = make_views(labels, 128, seed=0)
    wide1 = torch.zeros(64, 256); wide1[:, :128] = z1
    wide2 = torch.zeros(64, 256); wide2[:, :128] = z2
    a = _run(z1, z2, target=labels.tolist(), gamma=5.0, mode="soft", precision="fp32")
    crit = spcl_b200.SelfPacedSupConLoss(weight_update="soft"); crit.set_gamma(5.0)
    loss = crit(wide1.cuda()[:, :128], wide2.cuda()[:, :128], target=labels.tolist())
    assert np.isclose(loss.item(), a["loss"], rtol=1e-6)
    crit16 = spcl_b200.SelfPacedSupConLoss(weight_update="soft", validate=False); crit16.set_gamma(5.0)
    h1 = z1.cuda().half().requires_grad_(True)
    loss16 = crit16(h1, z2.cuda().half(), target=labels.tolist())
    loss16.backward()
    assert np.isclose(loss16.item(), a["loss"], rtol=2e-2) and h1.grad.dtype == torch.float16


def test_diagnostics_match_reference_semantics():
    labels = CFG1["labels_partition"].tolist()
    res = _run(CFG1["z1"], CFG1["z2"], target=labels, gamma=5.0, mode="soft", precision="fp32")
    crit = res["crit"]
    ref = dense_supcon(torch.from_numpy(CFG1["z1"]), torch.from_numpy(CFG1["z2"]), target=labels, gamma=5.0,
                       mode="soft")
    assert torch.equal(crit.pos_mask.cpu(), ref.pos_mask)            # bit-exact masks
    assert torch.equal(crit.neg_mask.cpu(), ref.neg_mask)
    assert torch.allclose(crit.sim_logits.cpu(), ref.sim_logits, atol=1e-4)
    assert torch.allclose(crit.sim_exp.cpu(), ref.sim_exp, atol=1e-5)
    assert torch.allclose(crit.sp_mask.cpu(), ref.sp_mask, atol=1e-4)
    assert isinstance(crit.downgrade_ratio, float) and crit.age_param == 5.0


def test_grad_scales_with_upstream_gradient():
    labels = acdc_meta_labels(64)["partition"]
    z1, z2 = make_views(labels, 128, seed=0)
    for precision in ("fp32", "bf16"):
        crit = spcl_b200.SupConLoss1(precision=precision)
        a = z1.cuda().requires_grad_(True)
        (crit(a, z2.cuda(), target=labels.tolist()) * 2.5).backward()
        g1 = a.grad.clone(); a.grad = None
        crit(a, z2.cuda(), target=labels.tolist()).backward()
        assert torch.allclose(g1, 2.5 * a.grad, rtol=1e-5, atol=1e-9)


def test_opcheck_registration():
    labels = acdc_meta_labels(64)["partition"].int().cuda()
    z1, z2 = make_views(labels.cpu(), 128, seed=0)
    args = (z1.cuda().requires_grad_(True), z2.cuda().requires_grad_(True), labels, None, 0.07, 5.0, 2, False, False)
    torch.library.opcheck(spcl_b200.ops.supcon_fwd, args, test_utils=("test_schema", "test_faketensor"))


# ------------------------------------------------------------------------------------------------
# projector tail
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(64, 256), (4096, 128), (32, 128, 32, 32), (7, 33), (3, 5, 7), (2, 256, 10, 10)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_l2norm_matches_torch(shape, dtype):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(*shape, generator=g).to(dtype).cuda().requires_grad_(True)
    y = spcl_b200.Normalize(dim=1)(x)
    go = torch.randn(*shape, generator=g).to(dtype).cuda()
    y.backward(go)
    x2 = x.detach().float().requires_grad_(True)
    y2 = F.normalize(x2, p=2, dim=1)
    y2.backward(go.float())
    tol = 1e-6 if dtype == torch.float32 else 1.6e-2
    assert (y.float() - y2).abs().max().item() <= tol
    assert (x.grad.float() - x2.grad).abs().max().item() <= tol * max(1.0, x2.grad.abs().max().item())


def test_l2norm_zero_row_uses_eps_like_torch():
    x = torch.zeros(4, 16).cuda(); x[1] = 1.0
    y = spcl_b200.normalize(x)
    assert torch.equal(y[0], torch.zeros(16).cuda()) and torch.allclose(y, F.normalize(x))


# ------------------------------------------------------------------------------------------------
# host-buffer front end (bench.py's e2e leg)
# ------------------------------------------------------------------------------------------------
def test_hostfeed_matches_direct_calls():
    """Staged (copy-stream) batches give the same losses / gradients as plain .cuda() calls, in order."""
    n, d, steps = 640, 128, 5
    labels = acdc_meta_labels(n)["patient"].to(torch.int32)
    batches = [make_views(labels, d, sigma=0.7, seed=s) for s in range(steps)]
    crit = spcl_b200.SelfPacedSupConLoss(weight_update="soft", precision="bf16", check_nan=False, validate=False)
    crit.set_gamma(6.0)
    want, want_g = [], []
    for z1, z2 in batches:
        a = z1.cuda().requires_grad_(True)
        b = z2.cuda().requires_grad_(True)
        loss = crit(a, b, target=labels.cuda())
        loss.backward()
        want.append(loss.item())
        want_g.append(a.grad.clone())
    feed = spcl_b200.HostFeed(n, d, "cuda", depth=2)
    pinned = [(z1.pin_memory(), z2.pin_memory()) for z1, z2 in batches]
    lab_h = labels.pin_memory()
    got_g = []
    feed.push(*pinned[0], lab_h)
    for k in range(steps):
        if k + 1 < steps:
            feed.push(*pinned[k + 1], lab_h)
        a, b, lab, slot = feed.pop()
        loss = crit(a, b, target=lab)
        loss.backward()
        got_g.append(a.grad)
        feed.release(slot, loss)
    got = feed.losses()
    assert feed.h2d_bytes == steps * (2 * n * d * 4 + n * 4)
    np.testing.assert_allclose(got, want, rtol=1e-5)
    for g, w in zip(got_g, want_g):
        # row sums and dZ are accumulated with atomics: same terms, different order; a last-bit change of a row
        # statistic can flip the bf16 rounding of a T element (2^-9 of that element)
        # (measured run to run: up to 3e-4 of max|g|; anything structural -- a stale slot, a wrong batch -- is O(1))
        assert (g - w).abs().max().item() <= 1e-3 * w.abs().max().item()
    with pytest.raises(RuntimeError):
        for _ in range(3):
            feed.push(*pinned[0], lab_h)


# ---- symmetric pass A: split API == one call == rectangular pass --------------------------------
@pytest.mark.parametrize("n,d,kind,mode", [(1000, 128, "patient", "none"), (1000, 128, "partition", "soft"),
                                           (331, 256, "cycle", "soft"), (2048, 64, "self", "soft")])
@pytest.mark.parametrize("nparts", [1, 3])
def test_stats_parts_plus_finish_equals_fused_forward(n, d, kind, mode, nparts):
    """spcl_supcon_stats_part_bf16 x nparts (+ sum) + spcl_supcon_fwd_finish_bf16 == spcl_supcon_fwd_bf16, and the
    symmetric whole-launch pass == the rectangular pass a row shard runs (two row ranges)."""
    from spcl_b200.ops import _ptr, _stream, pad_to
    labels = torch.arange(n).int() if kind == "self" else acdc_meta_labels(n)[kind].int()
    z1, z2 = make_views(labels, d, sigma=0.7, seed=3)
    z1, z2, lab = z1.cuda(), z2.cuda(), labels.cuda()
    N, n_pad, d_pad = 2 * n, pad_to(2 * n, nat.TILE), pad_to(d, 64)
    dev = z1.device
    st = _stream(z1)
    zpack = torch.empty(n_pad, d_pad, dtype=torch.bfloat16, device=dev)
    labels_full = torch.empty(n_pad, dtype=torch.int32, device=dev)
    sig = torch.empty(n_pad // nat.TILE, 4, dtype=torch.int32, device=dev)
    scratch = torch.empty(3, dtype=torch.float32, device=dev)
    nat.call("spcl_supcon_prepare_bf16", _ptr(z1), _ptr(z2), n, d, z1.stride(0), z2.stride(0), _ptr(lab),
             _ptr(zpack), n_pad, d_pad, _ptr(labels_full), _ptr(sig), _ptr(scratch), st)
    inv_tau, gamma, m = 1.0 / 0.07, 6.0, MODE[mode]

    def fused(ranges):
        acc = torch.empty(n_pad, 4, dtype=torch.float32, device=dev)
        rs = torch.zeros(4, n_pad, dtype=torch.float32, device=dev)
        pt = torch.zeros(3, dtype=torch.float32, device=dev)
        for rb, re in ranges:
            nat.call("spcl_supcon_fwd_bf16", _ptr(zpack), N, n_pad, d_pad, _ptr(labels_full), _ptr(sig), rb, re,
                     inv_tau, gamma, m, _ptr(acc), _ptr(rs), _ptr(pt), st)
        return rs[:, :N].cpu().numpy().astype(np.float64), pt.cpu().numpy().astype(np.float64)

    rs_one, pt_one = fused([(0, N)])                                  # symmetric pass
    cut = 128 * max(1, (N // 128) // 3)
    rs_rect, pt_rect = fused([(0, cut), (cut, N)])                    # two row shards: rectangular pass

    acc = torch.zeros(n_pad, 4, dtype=torch.float32, device=dev)
    for part in range(nparts):                                        # what the ranks do, then all-reduce
        nat.call("spcl_supcon_stats_part_bf16", _ptr(zpack), N, n_pad, d_pad, _ptr(labels_full), _ptr(sig), part,
                 nparts, inv_tau, m, _ptr(acc), st)
    rs = torch.zeros(4, n_pad, dtype=torch.float32, device=dev)
    pt = torch.zeros(3, dtype=torch.float32, device=dev)
    nat.call("spcl_supcon_fwd_finish_bf16", _ptr(zpack), N, n_pad, d_pad, _ptr(labels_full), _ptr(sig), 0, N,
             inv_tau, gamma, m, _ptr(acc), _ptr(rs), _ptr(pt), st)
    rs_split, pt_split = rs[:, :N].cpu().numpy().astype(np.float64), pt.cpu().numpy().astype(np.float64)

    for other_rs, other_pt in ((rs_rect, pt_rect), (rs_split, pt_split)):
        np.testing.assert_array_equal(other_rs[1], rs_one[1])                      # 1 / c_i: counts are exact
        np.testing.assert_allclose(other_rs[0], rs_one[0], rtol=0, atol=2e-5)      # logD: summation order only
        np.testing.assert_allclose(other_rs[2], rs_one[2], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(other_rs[3], rs_one[3], rtol=1e-4)
        np.testing.assert_allclose(other_pt, pt_one, rtol=2e-5)


# ---- eager direct route / CUDA-graph replay == registered custom ops ----------------------------
@pytest.mark.parametrize("n,d,mode,precision", [(64, 128, "soft", "fp32"), (256, 256, "hard", "fp32"),
                                                 (300, 128, "soft", "bf16")])
def test_cuda_graph_replay_matches_eager(n, d, mode, precision):
    labels = acdc_meta_labels(n)["patient"]
    outs = {}
    for graphed in (False, True):
        crit = spcl_b200.SelfPacedSupConLoss(weight_update=mode, correct_grad=True, precision=precision,
                                             cuda_graph=graphed)
        crit.set_gamma(5.0)
        res = []
        for seed in (0, 1, 2):                      # three different batches through the same module / graph
            z1, z2 = make_views(labels, d, sigma=0.7, seed=seed)
            a, b = z1.cuda().requires_grad_(True), z2.cuda().requires_grad_(True)
            loss = crit(a, b, target=labels.tolist())
            (3.0 * loss).backward()
            res.append((loss.item(), crit.downgrade_ratio, a.grad.clone(), b.grad.clone()))
        outs[graphed] = res
    for (l0, r0, ga0, gb0), (l1, r1, ga1, gb1) in zip(outs[False], outs[True]):
        # same kernels, same inputs: only the atomics' summation order differs
        assert np.isclose(l0, l1, rtol=1e-5) and np.isclose(r0, r1, rtol=1e-5)
        tol = 2e-5 * max(ga0.abs().max().item(), 1e-12) if precision == "fp32" else 2e-3 * ga0.abs().max().item()
        assert (ga0 - ga1).abs().max().item() <= tol and (gb0 - gb1).abs().max().item() <= tol


def test_direct_route_matches_custom_op():
    """The modules' eager route and the registered ``spcl::supcon_fwd`` op give the same loss and gradients."""
    from spcl_b200 import ops
    n, d = 192, 128
    labels = acdc_meta_labels(n)["cycle"].int().cuda()
    z1, z2 = make_views(labels.cpu(), d, sigma=0.7, seed=5)
    outs = []
    for direct in (False, True):
        a, b = z1.cuda().requires_grad_(True), z2.cuda().requires_grad_(True)
        if direct:
            scalars, _ = ops.supcon_fwd_eager(a, b, labels, None, 0.07, 4.0, nat.MODE_SOFT, True, False)
        else:
            scalars = ops.supcon_fwd(a, b, labels, None, 0.07, 4.0, nat.MODE_SOFT, True, False)[0]
        scalars[0].backward()
        outs.append((scalars[0].item(), a.grad, b.grad))
    assert np.isclose(outs[0][0], outs[1][0], rtol=1e-6)
    for k in (1, 2):                                  # same kernels: only the atomics' summation order differs
        assert (outs[0][k] - outs[1][k]).abs().max().item() <= 2e-5 * outs[0][k].abs().max().item()


# ---- fused projector tail ----------------------------------------------------------
@pytest.mark.parametrize("shape,mode", [((8, 128, 8, 8), "soft"), ((3, 96, 7, 5), "hard"), ((640, 128), "soft"),
                                        ((4, 256, 16, 16), "none"), ((40, 64), "soft")])
def test_forward_raw_equals_normalize_reshape_forward(shape, mode):
    """forward_raw(x1, x2) == forward(rows(F.normalize(x1)), rows(F.normalize(x2))): loss, ratio and the gradients
    w.r.t. the UN-normalised projector outputs (through F.normalize's own backward on the unfused side)."""
    g = torch.Generator().manual_seed(7)
    b, d = shape[0], shape[1]
    n = int(np.prod(shape)) // d
    n_cls = max(2, n // 6)
    labels = torch.randint(0, n_cls, (n,), generator=g)
    cent = torch.randn(n_cls, d, generator=g)
    def raw():
        rows = (cent[labels] + 0.7 * torch.randn(n, d, generator=g)) * (0.5 + torch.rand(n, 1, generator=g))
        return rows.reshape(b, -1, d).permute(0, 2, 1).reshape(shape).contiguous()
    x1, x2 = raw(), raw()
    def crit_():
        if mode == "none":
            return spcl_b200.SupConLoss1(precision="bf16")
        c = spcl_b200.SelfPacedSupConLoss(weight_update=mode, correct_grad=True, precision="bf16")
        c.set_gamma(6.0)
        return c
    rows_of = lambda y: y.reshape(b, d, -1).permute(0, 2, 1).reshape(n, d)
    a0, b0 = x1.cuda().requires_grad_(True), x2.cuda().requires_grad_(True)
    c0 = crit_()
    l0 = c0(rows_of(F.normalize(a0, dim=1)), rows_of(F.normalize(b0, dim=1)), target=labels.tolist())
    l0.backward()
    a1, b1 = x1.cuda().requires_grad_(True), x2.cuda().requires_grad_(True)
    c1 = crit_()
    l1 = c1.forward_raw(a1, b1, target=labels.tolist())
    l1.backward()
    assert np.isclose(l0.item(), l1.item(), rtol=2e-5), (l0.item(), l1.item())
    if mode != "none":
        assert np.isclose(c0.downgrade_ratio, c1.downgrade_ratio, rtol=1e-4)
    for u, w in ((a0.grad, a1.grad), (b0.grad, b1.grad)):
        assert w.shape == u.shape
        assert (u - w).abs().max().item() <= 2e-3 * u.abs().max().item()      # dZ atomics order, same formula


# ---- oracle-based (not self-referential) checks of the f1 / f3 entry points ---------------------------------------
@pytest.mark.parametrize("shape,mode,gamma", [((16, 128, 8, 8), "soft", 6.0), ((6, 96, 7, 5), "hard", 5.0),
                                              ((1280, 128), "soft", 4.0), ((4, 256, 16, 16), "none", 1e6)])
def test_forward_raw_against_the_oracle(shape, mode, gamma):
    """``forward_raw`` (normalise + reshape + concat + bf16 pack fused into the loss) against the fp64
    closed form: the oracle is fed the fp64-normalised rows rounded to bf16 (what the tensor-core kernels consume) and
    its gradient rows are pulled back through the normalisation in numpy (nn.py:35-36:
    gx = inv (g - y <y, g>), comparable.py:398-404 for the row order)."""
    g = torch.Generator().manual_seed(11)
    b, d = shape[0], shape[1]
    n = int(np.prod(shape)) // d
    inner = n // b
    n_cls = max(2, n // 8)
    labels = torch.randint(0, n_cls, (n,), generator=g)
    cent = torch.randn(n_cls, d, generator=g)

    def raw():
        rows = (cent[labels] + 0.7 * torch.randn(n, d, generator=g)) * (0.5 + torch.rand(n, 1, generator=g))
        return rows.reshape(b, inner, d).permute(0, 2, 1).reshape(shape).contiguous(), rows
    (x1, r1), (x2, r2) = raw(), raw()
    if mode == "none":
        crit = spcl_b200.SupConLoss1(precision="bf16")
    else:
        crit = spcl_b200.SelfPacedSupConLoss(weight_update=mode, correct_grad=True, precision="bf16")
        crit.set_gamma(gamma)
    a, bb = x1.cuda().requires_grad_(True), x2.cuda().requires_grad_(True)
    loss = crit.forward_raw(a, bb, target=labels.tolist())
    loss.backward()

    def unit_rows(r):                                            # fp64 normalise, then the kernels' bf16 rounding
        r = r.double()
        inv = 1.0 / r.norm(dim=1, keepdim=True).clamp_min(1e-12)
        y = r * inv
        return y, inv, y.float().bfloat16().double()
    y1, inv1, q1 = unit_rows(r1)
    y2, inv2, q2 = unit_rows(r2)
    ref = supcon_closed_form(q1.numpy(), q2.numpy(), target=labels.tolist(), gamma=gamma, mode=mode,
                             correct_grad=(mode != "none"))
    assert np.isclose(loss.item(), ref["loss"], rtol=BF16_TIGHT_LOSS_RTOL), (loss.item(), ref["loss"])
    if mode != "none":
        assert np.isclose(crit.downgrade_ratio, ref["ratio"], rtol=2e-4)
    for got, dz, y, inv in ((a.grad, ref["dz1"], y1, inv1), (bb.grad, ref["dz2"], y2, inv2)):
        dz = torch.from_numpy(np.asarray(dz)).double()
        gx_rows = inv * (dz - y * (y * dz).sum(1, keepdim=True))
        want = gx_rows.reshape(b, inner, d).permute(0, 2, 1).reshape(shape)
        got = got.double().cpu()
        rel = (got - want).abs().max().item() / want.abs().max().item()
        cos = (got * want).sum().item() / (got.norm().item() * want.norm().item())
        assert rel < 3e-2 and cos > BF16_TIGHT_COS, (rel, cos)


def test_grouped_forward_against_the_oracle():
    """The grouped launch (K meta-label problems per kernel launch) against the fp64 closed form of
    every problem, at the fp32 path's tolerances -- including a ragged group (different n and d per problem)."""
    meta = acdc_meta_labels(256)
    specs = [("partition", 256, 256, "soft", 5.0, True), ("patient", 256, 256, "hard", 3.5, False),
             ("cycle", 200, 128, "soft", 2.0, True)]
    crits, feats, targets, refs = [], [], [], []
    for kind, n, d, mode, gamma, cg in specs:
        lab = meta[kind][:n]
        z1, z2 = make_views(lab, d, sigma=0.7, seed=3)
        c = spcl_b200.SelfPacedSupConLoss(weight_update=mode, correct_grad=cg, precision="fp32")
        c.set_gamma(gamma)
        crits.append(c)
        feats.append((z1.cuda().requires_grad_(True), z2.cuda().requires_grad_(True)))
        targets.append(lab.tolist())
        refs.append(supcon_closed_form(z1.numpy(), z2.numpy(), target=lab.tolist(), gamma=gamma, mode=mode,
                                       correct_grad=cg))
    for graph in (False, True):
        for a, b in feats:
            a.grad = b.grad = None
        losses = spcl_b200.grouped_forward(crits, feats, targets, cuda_graph=graph)
        sum(losses).backward()
        for c, (a, b), l, ref in zip(crits, feats, losses, refs):
            assert np.isclose(l.item(), ref["loss"], rtol=FP32_LOSS_RTOL), (graph, l.item(), ref["loss"])
            assert np.isclose(c.downgrade_ratio, ref["ratio"], rtol=1e-4)
            rel, cos = _grad_metrics(dict(dz1=a.grad.cpu().numpy(), dz2=b.grad.cpu().numpy()), ref)
            assert rel < FP32_GRAD_REL, (graph, rel, cos)


def test_gamma_zero_gives_zero_loss_like_the_reference():
    """PScheduler's default begin_value is 0 (infonce.py:34-53): the reference then weights every positive with 0."""
    labels = acdc_meta_labels(96)["patient"]
    z1, z2 = make_views(labels, 64, sigma=0.7, seed=2)
    for precision, n_rep in (("fp32", 1), ("bf16", 1)):
        for mode in ("soft", "hard"):
            res = _run(z1, z2, target=labels.tolist(), gamma=0.0, mode=mode, precision=precision)
            assert res["loss"] == 0.0 and res["ratio"] == 0.0
            assert np.abs(res["dz1"]).max() == 0.0 and np.abs(res["dz2"]).max() == 0.0


def test_low_temperature_falls_back_to_the_fp32_kernels_under_auto():
    """1 / T > 40 is outside the tensor-core kernels' range: precision="auto" must still compute the loss."""
    labels = acdc_meta_labels(600)["partition"]
    z1, z2 = make_views(labels, 64, sigma=0.7, seed=4)
    res = _run(z1, z2, target=labels.tolist(), gamma=30.0, mode="soft", temperature=0.02, precision="auto")
    ref = supcon_closed_form(z1.numpy(), z2.numpy(), target=labels.tolist(), gamma=30.0, mode="soft", temperature=0.02)
    assert np.isclose(res["loss"], ref["loss"], rtol=1e-4), (res["loss"], ref["loss"])
    with pytest.raises(nat.SpclError):
        _run(z1, z2, target=labels.tolist(), gamma=30.0, mode="soft", temperature=0.02, precision="bf16")


@pytest.mark.parametrize("mode,gamma,cg", [("none", 1e6, False), ("soft", 6.0, True), ("hard", 5.5, False)])
def test_all_positive_tiles_against_the_oracle(mode, gamma, cg):
    """Label runs of 256 anchors: every 128 x 128 tile is either all-positive or all-negative, so the sp pass and the
    backward take their packed all-positive paths (what cfg3's slice labels run at full size); a second labelling with
    runs of 96 mixes them with the generic per-pair path inside one problem.  fp64 oracle on the bf16-rounded operands."""
    for run in (256, 96):
        n, d = 1024, 128
        labels = (torch.arange(n) // run)
        z1, z2 = make_views(labels, d, sigma=0.7, seed=9)
        z1, z2 = z1.bfloat16().float(), z2.bfloat16().float()
        cls = "SupConLoss1" if mode == "none" else "SP"
        res = _run(z1, z2, cls=cls, target=labels.int().numpy(), gamma=gamma, mode=mode, correct_grad=cg,
                   precision="bf16", validate=False)
        ref = supcon_closed_form(z1.numpy(), z2.numpy(), target=labels.tolist(), gamma=gamma, mode=mode, correct_grad=cg)
        assert np.isclose(res["loss"], ref["loss"], rtol=BF16_TIGHT_LOSS_RTOL), (run, res["loss"], ref["loss"])
        if mode != "none":
            assert np.isclose(res["ratio"], ref["ratio"], rtol=3e-4), (run, res["ratio"], ref["ratio"])
        rel, cos = _grad_metrics(res, ref)
        assert rel < BF16_TIGHT_GRAD_REL and cos > BF16_TIGHT_COS, (run, rel, cos)


@pytest.mark.parametrize("flag", [16384, 32768], ids=["wide", "cta_pair"])
@pytest.mark.parametrize("n,run,mode,gamma", [(1024, 256, "soft", 6.0), (1024, 96, "hard", 5.5), (1536, 1, "none", 1e6),
                                              (4096, 1, "soft", 8.0)])
def test_optional_backward_kernels_against_the_oracle(n, run, mode, gamma, flag):
    """The two opt-in backward kernels against the fp64 oracle, same tolerances as the default backward:
    `bwd_wide_kernel` (128 x 256 S tiles: SPCL_BWD_WIDE=1 / debug flag 16384; profiles/r02zb_bwd_wide_experiment.txt) and
    `bwd2_kernel` (a CTA pair per 256 anchor rows, `tcgen05 cta_group::2`: SPCL_PAIR=1 / debug flag 32768).
    All-positive, mixed and SimCLR labellings; 2 n / 128 column tiles = 16 ... 64 (pairs that wrap the wide kernel's
    five-tile slot ring included)."""
    import ctypes
    h = nat.lib()
    h.spcl_debug_set_flags.argtypes = [ctypes.c_int]
    d = 128
    labels = (torch.arange(n) // run)
    z1, z2 = make_views(labels, d, sigma=0.7, seed=11)
    z1, z2 = z1.bfloat16().float(), z2.bfloat16().float()
    cls = "SupConLoss1" if mode == "none" else "SP"
    h.spcl_debug_set_flags(flag)
    try:
        res = _run(z1, z2, cls=cls, target=labels.int().numpy(), gamma=gamma, mode=mode, precision="bf16", validate=False)
    finally:
        h.spcl_debug_set_flags(0)
    base = _run(z1, z2, cls=cls, target=labels.int().numpy(), gamma=gamma, mode=mode, precision="bf16", validate=False)
    ref = supcon_closed_form(z1.numpy(), z2.numpy(), target=labels.tolist(), gamma=gamma, mode=mode)
    rel, cos = _grad_metrics(res, ref)
    assert rel < BF16_TIGHT_GRAD_REL and cos > BF16_TIGHT_COS, (rel, cos)
    # the backward kernels evaluate the same T and differ in the accumulation order of S and dZ only; under the hard
    # rule a pair within rounding of gamma may take the other weight in the other kernel, so only the oracle bound applies
    if mode != "hard":
        g, b = np.concatenate([res["dz1"], res["dz2"]]), np.concatenate([base["dz1"], base["dz2"]])
        assert np.abs(g - b).max() <= 2e-3 * np.abs(b).max(), np.abs(g - b).max() / np.abs(b).max()


# ------------------------------------------------------------------------------------------------
# the two fp32 routes of the reference's own batch sizes: one cooperative launch (default) / one launch per stage
# ------------------------------------------------------------------------------------------------
def test_fused_small_batch_route_is_taken_where_it_fits():
    from spcl_b200 import ops
    cap = ops.fused_capacity("cuda")
    assert cap >= 148, cap                                       # at least one CTA per SM on a B200
    assert ops.fused_fits([(256, 256)] * 3, "cuda")              # cfg2: 3 x 64 tiles
    assert not ops.fused_fits([(2048, 128)], "cuda")             # 64 x 64 tiles: the multi-launch kernels take it
    assert not ops.fused_fits([(64, 128)] * 3 + [(4096, 64)], "cuda")


@pytest.mark.parametrize("fused", ["1", "0"])
@pytest.mark.parametrize("name", ["sp_soft_g5_cg0_patient", "sp_hard_g2_cg1_partition", "supcon1_partition",
                                  "sp_soft_g5_cg0_simclr_none", "sp_soft_g5_cg0_tensor_target",
                                  "sp_soft_g5_cg0_partition_t0.2", "sp_hard_g5_cg0_composite"])
def test_fp32_routes_match_reference_cfg1(name, fused, monkeypatch):
    """Both routes against the goldens of the unmodified reference (the route is chosen per call from the
    environment); the case names are a subset of CFG1.cases."""
    assert name in CFG1.cases
    monkeypatch.setenv("SPCL_FUSED_SMALL", fused)
    kw = parse_cfg1_case(name, CFG1)
    res = _run(CFG1["z1"], CFG1["z2"], precision="fp32", **kw)
    ref = CFG1.case(name)
    assert np.isclose(res["loss"], ref["loss"], rtol=FP32_LOSS_RTOL), (res["loss"], ref["loss"])
    if kw["cls"] != "SupConLoss1":
        assert np.isclose(res["ratio"], ref["ratio"], rtol=1e-5, atol=1e-7)
    assert _grad_metrics(res, ref)[0] < FP32_GRAD_REL


@pytest.mark.parametrize("name", CFG2.cases)
def test_multi_launch_fp32_route_matches_reference_cfg2(name, monkeypatch):
    monkeypatch.setenv("SPCL_FUSED_SMALL", "0")
    z1, z2 = CFG2[f"{name}/z1"], CFG2[f"{name}/z2"]
    res = _run(z1, z2, target=CFG2[f"{name}/labels"].tolist(), gamma=float(CFG2[f"{name}/gamma"]), mode="soft",
               precision="fp32")
    ref = CFG2.case(name)
    assert np.isclose(res["loss"], ref["loss"], rtol=FP32_LOSS_RTOL)
    assert _grad_metrics(res, ref)[0] < FP32_GRAD_REL


@pytest.mark.parametrize("n,d,mode,gamma", [(5, 16, "soft", 3.0), (97, 200, "hard", 4.0), (300, 256, "none", 1e6),
                                            (544, 96, "soft", 6.0)])
def test_fused_small_batch_against_the_oracle(n, d, mode, gamma):
    """Ragged sizes (N not a multiple of the 64-anchor tile, d not a multiple of 16 / above 128: two column passes of
    the dZ stage), up to the largest single problem the cooperative grid holds (17 x 17 tiles)."""
    from spcl_b200 import ops
    assert ops.fused_fits([(n, d)], "cuda")
    labels = acdc_meta_labels(n)["patient"] if n > 8 else torch.tensor([0, 1, 0, 1, 0])
    z1, z2 = make_views(labels, d, sigma=0.7, seed=3)
    cls = "SupConLoss1" if mode == "none" else "SP"
    res = _run(z1, z2, cls=cls, target=labels.tolist(), gamma=gamma, mode=mode, correct_grad=True, precision="fp32")
    ref = supcon_closed_form(z1.numpy(), z2.numpy(), target=labels.tolist(), gamma=gamma, mode=mode, correct_grad=True)
    assert np.isclose(res["loss"], ref["loss"], rtol=FP32_LOSS_RTOL), (res["loss"], ref["loss"])
    if mode != "none":
        assert np.isclose(res["ratio"], ref["ratio"], rtol=1e-5, atol=1e-7)
    assert _grad_metrics(res, ref)[0] < FP32_GRAD_REL


def test_declined_cooperative_launch_falls_back_to_the_staged_route(monkeypatch):
    """`spcl_supcon_group_fused_f32` may decline (SPCL_ERR_UNSUPPORTED: the cooperative grid cannot be resident, e.g. on an
    SM-partitioned context); single calls and grouped calls must then produce the same results on the staged kernels."""
    from spcl_b200 import ops
    declined = []

    def decline(probs, count, st, device):
        declined.append(count)
        return False
    monkeypatch.setattr(ops, "_fused_launch", decline)
    name = CFG2.cases[0]
    z1, z2 = CFG2[f"{name}/z1"], CFG2[f"{name}/z2"]
    labels, gamma = CFG2[f"{name}/labels"].tolist(), float(CFG2[f"{name}/gamma"])
    res = _run(z1, z2, target=labels, gamma=gamma, mode="soft", precision="fp32")
    ref = CFG2.case(name)
    assert declined == [1]
    assert np.isclose(res["loss"], ref["loss"], rtol=FP32_LOSS_RTOL)
    assert _grad_metrics(res, ref)[0] < FP32_GRAD_REL
    crit = spcl_b200.SelfPacedSupConLoss(weight_update="soft", precision="fp32")
    crit.set_gamma(gamma)
    a, b = torch.as_tensor(z1).cuda().requires_grad_(True), torch.as_tensor(z2).cuda().requires_grad_(True)
    (loss,) = spcl_b200.grouped_forward([crit], [(a, b)], [labels])
    loss.backward()
    assert declined == [1, 1]
    assert np.isclose(loss.item(), ref["loss"], rtol=FP32_LOSS_RTOL)
    got = dict(dz1=a.grad.cpu().numpy(), dz2=b.grad.cpu().numpy())
    assert _grad_metrics(got, ref)[0] < FP32_GRAD_REL
