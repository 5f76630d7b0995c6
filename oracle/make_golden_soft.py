"""Generate tests/golden/soft_weight_cases.npz by running the UNMODIFIED ``contrastyou/losses/contrast_loss.py``
(SupConLoss2 / SupConLoss3 / SupConLoss4).  TEST INFRASTRUCTURE; build container only (needs /root/reference):

    python oracle/make_golden_soft.py

The file imports ``matplotlib.pyplot`` and ``deepclustering2.writer.SummaryWriter`` for its figure hook only; both are
replaced by stubs (the arithmetic never touches them).  Every case stores inputs, loss and autograd gradients.
"""
from __future__ import annotations

import importlib.util
import pathlib
import sys
import types

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
REF_FILE = pathlib.Path("/root/reference/contrastyou/losses/contrast_loss.py")
OUT = ROOT / "tests" / "golden" / "soft_weight_cases.npz"


def load_reference():
    mpl = types.ModuleType("matplotlib")
    mpl.get_backend = lambda: "agg"
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    for name in ("deepclustering2", "deepclustering2.writer"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["deepclustering2.writer"].SummaryWriter = object
    spec = importlib.util.spec_from_file_location("ref_contrast_loss", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run(crit, z1, z2, **kw):
    a = z1.clone().requires_grad_(True)
    b = z2.clone().requires_grad_(True)
    loss = crit(proj_feat1=a, proj_feat2=b, **kw)
    loss.backward()
    return dict(loss=np.float64(loss.item()), dz1=a.grad.numpy().copy(), dz2=b.grad.numpy().copy())


def main():
    ref = load_reference()
    g = torch.Generator().manual_seed(0)
    out, names = {}, []

    def views(n, d, n_cls):
        lab = torch.randint(0, n_cls, (n,), generator=g)
        cent = torch.randn(n_cls, d, generator=g)
        z = [torch.nn.functional.normalize(cent[lab] + 0.7 * torch.randn(n, d, generator=g), dim=1) for _ in range(2)]
        return lab, z[0], z[1]

    def soft_weight(lab, n):
        # a "softened" positive matrix: 1 on equal labels, a decaying weight on neighbouring classes, 0 elsewhere
        diff = (lab[:, None] - lab[None, :]).abs().float()
        return torch.where(diff == 0, torch.ones(n, n), torch.where(diff == 1, torch.full((n, n), 0.35), torch.zeros(n, n)))

    def add(name, res, **inputs):
        names.append(name)
        for k, v in {**inputs, **res}.items():
            out[f"{name}/{k}"] = np.asarray(v)

    for n, d, n_cls in ((24, 32, 4), (70, 128, 6)):
        lab, z1, z2 = views(n, d, n_cls)
        for out_mode in (True, False):
            tag = "out" if out_mode else "in"
            add(f"loss2_{tag}_target_n{n}", run(ref.SupConLoss2(temperature=0.07, out_mode=out_mode), z1, z2, target=lab.tolist()),
                z1=z1, z2=z2, target=lab)
            tri = torch.randint(0, 3, (n, n), generator=g).float()
            tri[torch.arange(n), torch.arange(n)] = 1.0
            tri[:, 0] = 1.0                                         # every row keeps a positive
            tri[0, :] = 1.0
            add(f"loss2_{tag}_mask_n{n}", run(ref.SupConLoss2(temperature=0.1, out_mode=out_mode), z1, z2, mask=tri),
                z1=z1, z2=z2, mask=tri, temperature=0.1)
            pw = soft_weight(lab, n)
            add(f"loss3_{tag}_n{n}", run(ref.SupConLoss3(temperature=0.07, out_mode=out_mode), z1, z2, pos_weight=pw),
                z1=z1, z2=z2, pos_weight=pw)
            w11, w22, w12 = soft_weight(lab, n), soft_weight(lab, n) * 0.8, soft_weight(lab, n) * 0.6
            add(f"loss4_{tag}_all_n{n}", run(ref.SupConLoss4(temperature=0.07, out_mode=out_mode), z1, z2,
                                              one2one_weight=w11, two2two_weight=w22, one2two_weight=w12),
                z1=z1, z2=z2, one2one=w11, two2two=w22, one2two=w12)
            # (two2two alone leaves the view-1 rows without any enabled pair: the reference raises on the NaN, :268-269)
            add(f"loss4_{tag}_no22_n{n}", run(ref.SupConLoss4(temperature=0.07, out_mode=out_mode), z1, z2,
                                               one2one_weight=w11, two2two_weight=None, one2two_weight=w12),
                z1=z1, z2=z2, one2one=w11, one2two=w12)
    out["names"] = np.asarray(names)
    OUT.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT} ({len(names)} cases)")


if __name__ == "__main__":
    main()
