"""Soft positive weights: SupConLoss2 / SupConLoss3 / SupConLoss4 of contrastyou/losses/contrast_loss.py.

CPU: the fp64 restatement (oracle/soft_weight.py) against outputs of the unmodified reference file
(tests/golden/soft_weight_cases.npz, oracle/make_golden_soft.py).  GPU: the CUDA kernels behind the modules of the
same names against the same goldens."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle.soft_weight import weighted_supcon, weights_loss2, weights_loss3, weights_loss4

G = np.load(GOLDEN / "soft_weight_cases.npz", allow_pickle=False)
NAMES = [str(n) for n in G["names"]]
# fp32 reference vs fp64 oracle / fp32 kernels: summation order only
LOSS_RTOL, GRAD_REL = 2e-5, 1e-4


def _get(name, key, default=None):
    k = f"{name}/{key}"
    return G[k] if k in G.files else default


def _case(name):
    z1, z2 = _get(name, "z1"), _get(name, "z2")
    temperature = float(_get(name, "temperature", 0.07))
    in_mode = "_in_" in name
    return z1, z2, temperature, in_mode


def _weights(name, n):
    if name.startswith("loss2"):
        t, m = _get(name, "target"), _get(name, "mask")
        return weights_loss2(n, target=None if t is None else t.tolist(), mask=m)
    if name.startswith("loss3"):
        return weights_loss3(_get(name, "pos_weight"))
    return weights_loss4(n, _get(name, "one2one"), _get(name, "two2two"), _get(name, "one2two"))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_the_reference(name):
    z1, z2, temperature, in_mode = _case(name)
    w, en = _weights(name, z1.shape[0])
    out = weighted_supcon(z1, z2, w, en, temperature=temperature, in_mode=in_mode)
    assert np.isclose(out["loss"], float(_get(name, "loss")), rtol=LOSS_RTOL), (out["loss"], float(_get(name, "loss")))
    for k in ("dz1", "dz2"):
        ref = _get(name, k)
        assert np.abs(out[k] - ref).max() <= GRAD_REL * np.abs(ref).max(), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_modules_match_the_reference(name):
    import spcl_b200
    z1, z2, temperature, in_mode = _case(name)
    a = torch.from_numpy(z1).cuda().requires_grad_(True)
    b = torch.from_numpy(z2).cuda().requires_grad_(True)
    cls = {"loss2": spcl_b200.SupConLoss2, "loss3": spcl_b200.SupConLoss3, "loss4": spcl_b200.SupConLoss4}[name[:5]]
    crit = cls(temperature=temperature, out_mode=not in_mode)
    t = lambda key: None if _get(name, key) is None else torch.from_numpy(_get(name, key)).cuda()
    if name.startswith("loss2"):
        tg = _get(name, "target")
        loss = crit(a, b, target=None if tg is None else tg.tolist(), mask=t("mask"))
    elif name.startswith("loss3"):
        loss = crit(a, b, pos_weight=t("pos_weight"))
    else:
        loss = crit(proj_feat1=a, proj_feat2=b, one2one_weight=t("one2one"), two2two_weight=t("two2two"),
                    one2two_weight=t("one2two"))
    loss.backward()
    assert np.isclose(loss.item(), float(_get(name, "loss")), rtol=LOSS_RTOL), (loss.item(), float(_get(name, "loss")))
    for got, k in ((a.grad, "dz1"), (b.grad, "dz2")):
        ref = _get(name, k)
        assert np.abs(got.cpu().numpy() - ref).max() <= GRAD_REL * np.abs(ref).max(), k
    assert crit.sim_exp.shape == (2 * z1.shape[0],) * 2


@pytest.mark.gpu
def test_loss4_with_only_the_view2_block_raises_like_the_reference():
    """two2two_weight alone leaves every view-1 anchor without an enabled pair: 0/0 -> NaN -> RuntimeError (:268-269)."""
    import spcl_b200
    g = torch.Generator().manual_seed(0)
    z = [torch.nn.functional.normalize(torch.randn(16, 32, generator=g), dim=1).cuda() for _ in range(2)]
    w = torch.ones(16, 16).cuda()
    with pytest.raises(RuntimeError):
        spcl_b200.SupConLoss4()(proj_feat1=z[0], proj_feat2=z[1], two2two_weight=w)


def test_weighted_modules_need_cuda():
    import spcl_b200
    z = torch.nn.functional.normalize(torch.randn(8, 16), dim=1)
    with pytest.raises(RuntimeError):
        spcl_b200.SupConLoss3()(z, z.clone(), pos_weight=torch.eye(8))
