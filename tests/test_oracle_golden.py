"""The oracle restatements against outputs of the UNMODIFIED reference (tests/golden).

CPU only.  This is what pins the oracle (the reference ships no
golden vectors, so they are generated here by oracle/make_golden.py).
"""
import numpy as np
import pytest
import torch

from oracle.closed_form import codes_from_target, supcon_closed_form
from oracle.dense_port import dense_supcon
from conftest import Golden, excl_case_inputs, parse_cfg1_case

CFG1 = Golden("cfg1_n64_d128.npz")
TINY = Golden("tiny_n5_d16.npz")
CFG2 = Golden("cfg2_n256_d256.npz")
EXCL = Golden("excl_cases.npz")

# fp64 closed form vs the fp32 reference: the reference's own rounding is the floor.
LOSS_RTOL = 2e-5
GRAD_ATOL_REL = 2e-5   # max|diff| <= this * max|grad_ref|


def _check(res, ref, loss_rtol=LOSS_RTOL, grad_rel=GRAD_ATOL_REL):
    assert np.isclose(res["loss"], ref["loss"], rtol=loss_rtol, atol=1e-7), (res["loss"], ref["loss"])
    if not np.isnan(ref["ratio"]):
        assert np.isclose(res["ratio"], ref["ratio"], rtol=1e-5, atol=1e-7), (res["ratio"], ref["ratio"])
    for k in ("dz1", "dz2"):
        scale = np.abs(ref[k]).max()
        assert np.abs(res[k] - ref[k]).max() <= grad_rel * scale + 1e-9, k


@pytest.mark.parametrize("name", CFG1.cases)
def test_closed_form_cfg1(name):
    kw = parse_cfg1_case(name, CFG1)
    kw.pop("cls")
    res = supcon_closed_form(CFG1["z1"], CFG1["z2"], **kw)
    _check(res, CFG1.case(name))


@pytest.mark.parametrize("name", CFG1.cases)
def test_dense_port_cfg1(name):
    kw = parse_cfg1_case(name, CFG1)
    kw.pop("cls")
    z1 = torch.from_numpy(CFG1["z1"]).requires_grad_(True)
    z2 = torch.from_numpy(CFG1["z2"]).requires_grad_(True)
    if kw["mask"] is not None:
        kw["mask"] = torch.from_numpy(kw["mask"])
    if isinstance(kw["target"], np.ndarray):
        kw["target"] = torch.from_numpy(kw["target"])
    out = dense_supcon(z1, z2, **kw)
    out.loss.backward()
    ref = CFG1.case(name)
    # same algorithm, same dtype: should agree to fp32 rounding
    res = dict(loss=out.loss.item(), ratio=out.ratio, dz1=z1.grad.numpy(), dz2=z2.grad.numpy())
    _check(res, ref, loss_rtol=1e-6, grad_rel=1e-5)


@pytest.mark.parametrize("name", TINY.cases)
def test_closed_form_tiny(name):
    labels = TINY["labels"].tolist()
    kw = {
        "sp_soft_g3": dict(target=labels, gamma=3.0, mode="soft"),
        "sp_hard_g3": dict(target=labels, gamma=3.0, mode="hard"),
        "supcon1": dict(target=labels, mode="none"),
        "sp_simclr": dict(gamma=4.0, mode="soft"),
    }[name]
    res = supcon_closed_form(TINY["z1"], TINY["z2"], **kw)
    _check(res, TINY.case(name))


@pytest.mark.parametrize("name", CFG2.cases)
def test_closed_form_cfg2(name):
    res = supcon_closed_form(CFG2[f"{name}/z1"], CFG2[f"{name}/z2"], target=CFG2[f"{name}/labels"].tolist(),
                             gamma=float(CFG2[f"{name}/gamma"]), mode="soft", block=128)
    _check(res, CFG2.case(name))


@pytest.mark.parametrize("name", EXCL.cases)
def test_closed_form_exclude_other_pos(name):
    z1, z2, kw = excl_case_inputs(EXCL, name)
    res = supcon_closed_form(z1, z2, mode="excl", **kw)
    _check(res, EXCL.case(name))


@pytest.mark.parametrize("name", EXCL.cases)
def test_dense_port_exclude_other_pos(name):
    z1, z2, kw = excl_case_inputs(EXCL, name)
    a = torch.from_numpy(z1).requires_grad_(True)
    b = torch.from_numpy(z2).requires_grad_(True)
    if kw["mask"] is not None:
        kw["mask"] = torch.from_numpy(kw["mask"])
    out = dense_supcon(a, b, mode="excl", **kw)
    out.loss.backward()
    res = dict(loss=out.loss.item(), ratio=float("nan"), dz1=a.grad.numpy(), dz2=b.grad.numpy())
    ref = dict(EXCL.case(name), ratio=np.float64("nan"))
    _check(res, ref, loss_rtol=1e-6, grad_rel=1e-5)


def test_identities():
    """Identities implied by the reference's __main__ blocks."""
    z1, z2 = CFG1["z1"], CFG1["z2"]
    t = CFG1["labels_partition"].tolist()
    a = supcon_closed_form(z1, z2, target=t, gamma=1e6, mode="hard")
    b = supcon_closed_form(z1, z2, target=t, mode="none")
    assert np.isclose(a["loss"], b["loss"], rtol=1e-12)          # gamma -> inf == SupConLoss1
    c = supcon_closed_form(z1, z2, target=list(range(len(t))), gamma=5.0, mode="soft")
    e = supcon_closed_form(z1, z2, gamma=5.0, mode="soft")
    assert np.isclose(c["loss"], e["loss"], rtol=1e-12)          # range(n) == no target == SimCLR
    np.testing.assert_allclose(c["dz1"], e["dz1"], rtol=1e-12, atol=1e-15)


def test_block_size_independent():
    z1, z2 = CFG1["z1"], CFG1["z2"]
    t = CFG1["labels_patient"].tolist()
    a = supcon_closed_form(z1, z2, target=t, gamma=4.0, mode="soft", block=7)
    b = supcon_closed_form(z1, z2, target=t, gamma=4.0, mode="soft", block=512)
    assert np.isclose(a["loss"], b["loss"], rtol=1e-13)
    np.testing.assert_allclose(a["dz2"], b["dz2"], rtol=1e-10, atol=1e-15)


def test_gradient_matches_finite_differences():
    rng = np.random.default_rng(0)
    n, d = 6, 8
    z1 = rng.normal(size=(n, d)); z1 /= np.linalg.norm(z1, axis=1, keepdims=True)
    z2 = rng.normal(size=(n, d)); z2 /= np.linalg.norm(z2, axis=1, keepdims=True)
    t = [0, 1, 0, 1, 2, 2]
    # mode "none": W is constant so finite differences see the same function the gradient describes
    base = supcon_closed_form(z1, z2, target=t, mode="none", temperature=0.5)
    eps = 1e-6
    for (i, k) in [(0, 0), (3, 5), (5, 2)]:
        zp = z1.copy(); zp[i, k] += eps
        zm = z1.copy(); zm[i, k] -= eps
        fd = (supcon_closed_form(zp, z2, target=t, mode="none", temperature=0.5, want_grad=False)["loss"]
              - supcon_closed_form(zm, z2, target=t, mode="none", temperature=0.5, want_grad=False)["loss"]) / (2 * eps)
        assert np.isclose(fd, base["dz1"][i, k], rtol=1e-5, atol=1e-8)


def test_label_codes_follow_fp32_semantics():
    # python-list labels go through float32 in the reference (contrast_loss3.py:135): 2**24 and 2**24+1 collide
    codes = codes_from_target([2 ** 24, 2 ** 24 + 1, 5])
    assert codes[0] == codes[1] != codes[2]
    # integer tensors are compared exactly
    codes = codes_from_target(np.array([2 ** 24, 2 ** 24 + 1, 5], dtype=np.int64))
    assert len(set(codes.tolist())) == 3


def test_zero_positive_row_is_nan():
    # an anchor whose tri-state row has no positives gives 0/0 -> NaN (reference raises RuntimeError, :203)
    n = 4
    rng = np.random.default_rng(1)
    z = rng.normal(size=(n, 8)); z /= np.linalg.norm(z, axis=1, keepdims=True)
    tri = np.zeros((n, n)); tri[1:, 1:] = 1.0
    res = supcon_closed_form(z, z[::-1].copy(), mask=tri, mode="none", want_grad=False)
    assert np.isnan(res["loss"])
    with pytest.raises(RuntimeError):
        dense_supcon(torch.from_numpy(z).float(), torch.from_numpy(z[::-1].copy()).float(),
                     mask=torch.from_numpy(tri).float(), mode="none")


# ---- dense front end: oracle and host-side point draw vs the reference-generated fixture --------
DENSE = Golden("dense_points.npz")


@pytest.mark.parametrize("name", DENSE.cases)
def test_dense_oracle_matches_reference_rows(name):
    from oracle.dense_frontend import dense_rows
    ph, pw, P, seed = (int(v) for v in DENSE[f"{name}/geom"])
    x = DENSE[f"{name}/x"]
    rows = dense_rows(x, ph, pw, points=DENSE[f"{name}/points"])
    np.testing.assert_allclose(rows, DENSE[f"{name}/rows"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(dense_rows(x, ph, pw), DENSE[f"{name}/all_rows"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("name", DENSE.cases)
def test_point_coordinates_reproduce_the_reference_draw(name):
    """bit-exact: same legacy numpy stream as ``with FixRandomSeed(seed): region_extractor(...)``."""
    import spcl_b200
    ph, pw, P, seed = (int(v) for v in DENSE[f"{name}/geom"])
    B = DENSE[f"{name}/x"].shape[0]
    pts = spcl_b200.point_coordinates(B, ph, pw, P, seed=seed)
    assert pts.dtype == torch.int32 and tuple(pts.shape) == (B, P)
    assert np.array_equal(pts.numpy(), DENSE[f"{name}/points"])
    # the global-generator form (seed=None) draws from numpy.random like the reference does
    np.random.seed(seed)
    assert np.array_equal(spcl_b200.point_coordinates(B, ph, pw, P).numpy(), DENSE[f"{name}/points"])


def test_dense_oracle_gradient_is_the_adjoint():
    from oracle.dense_frontend import dense_rows, dense_rows_grad
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 4, 7, 9))
    pts = np.array([[0, 5, 11], [3, 3, 8]])          # repeated point: gradients add
    for points in (None, pts):
        y0 = dense_rows(x, 3, 4, points)
        gy = rng.standard_normal(y0.shape)
        g = dense_rows_grad(x, 3, 4, gy, points)
        dx = rng.standard_normal(x.shape)
        h = 1e-6
        fd = ((dense_rows(x + h * dx, 3, 4, points) - dense_rows(x - h * dx, 3, 4, points)) * gy).sum() / (2 * h)
        assert np.isclose((g * dx).sum(), fd, rtol=1e-6, atol=1e-8)
