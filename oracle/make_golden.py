"""Generate tests/golden/*.npz by running the UNMODIFIED reference loss module.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

The reference file is loaded straight from /root/reference with
``importlib.util.spec_from_file_location`` (never ``import contrastyou`` -- its
``__init__`` creates directories next to the package).  ``matplotlib`` and
``deepclustering2`` are imported by the file but unused by the arithmetic
 and are replaced by empty stub modules.

Each case stores the inputs and the reference's loss / downgrade_ratio / autograd
gradients, so the tests never need the reference at run time.
"""
from __future__ import annotations

import importlib.util
import pathlib
import sys
import types

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
REF_FILE = pathlib.Path("/root/reference/contrastyou/losses/contrast_loss3.py")
OUT = ROOT / "tests" / "golden"


def load_reference():
    mpl = types.ModuleType("matplotlib")
    mpl.get_backend = lambda: "agg"
    mpl.use = lambda *a, **k: None
    sys.modules.setdefault("matplotlib", mpl)
    for name in ("deepclustering2", "deepclustering2.configparser", "deepclustering2.configparser._utils"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["deepclustering2.configparser._utils"].get_config = lambda *a, **k: {}
    from loguru import logger
    logger.disable("ref_contrast_loss3")
    spec = importlib.util.spec_from_file_location("ref_contrast_loss3", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _load_workloads():
    spec = importlib.util.spec_from_file_location(
        "spcl_workloads", ROOT / "self-paced-contrastive-learning_b200" / "workloads.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_reference(ref, z1, z2, *, cls, target=None, mask=None, gamma=None, mode="hard",
                  correct_grad=False, temperature=0.07):
    a = z1.clone().requires_grad_(True)
    b = z2.clone().requires_grad_(True)
    if cls == "SupConLoss1":
        crit = ref.SupConLoss1(temperature=temperature)
    elif cls == "SupConLoss1Excl":
        crit = ref.SupConLoss1(temperature=temperature, exclude_other_pos=True)
    else:
        crit = ref.SelfPacedSupConLoss(temperature=temperature, weight_update=mode, correct_grad=correct_grad)
        if gamma is not None:
            crit.set_gamma(gamma)
    kwargs = {}
    if mask is not None:
        kwargs["mask"] = mask
    elif target is not None:
        kwargs["target"] = target
    loss = crit(a, b, **kwargs)
    loss.backward()
    ratio = getattr(crit, "downgrade_ratio", float("nan"))
    return dict(loss=np.float64(loss.item()), ratio=np.float64(ratio),
                dz1=a.grad.numpy().copy(), dz2=b.grad.numpy().copy())


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref = load_reference()
    wl = _load_workloads()
    OUT.mkdir(parents=True, exist_ok=True)

    # ---------------- cfg1: n=64, d=128 ----------------
    n, d = 64, 128
    meta = wl.acdc_meta_labels(n)
    z1, z2 = wl.make_views(meta["partition"], d, sigma=0.7, seed=0)
    tri = torch.randint(0, 3, (n, n), generator=torch.Generator().manual_seed(7)).float()
    tri.fill_diagonal_(1.0)           # cross-view twin stays a positive -> c_i >= 1
    cases, store = [], {}
    store["z1"], store["z2"] = z1.numpy(), z2.numpy()
    store["tri_mask"] = tri.numpy()
    for key in ("partition", "patient", "cycle", "composite"):
        store["labels_" + key] = meta[key].numpy()

    def add(name, **kw):
        res = run_reference(ref, z1, z2, **kw)
        for k, v in res.items():
            store[f"{name}/{k}"] = v
        cases.append(name)

    for mode in ("hard", "soft"):
        for gamma in (2.0, 5.0, 20.0, 1e6):
            for cg in (False, True):
                add(f"sp_{mode}_g{gamma:g}_cg{int(cg)}_partition", cls="SP",
                    target=meta["partition"].tolist(), gamma=gamma, mode=mode, correct_grad=cg)
    for key in ("patient", "cycle", "composite"):
        add(f"sp_soft_g5_cg0_{key}", cls="SP", target=meta[key].tolist(), gamma=5.0, mode="soft")
        add(f"sp_hard_g5_cg0_{key}", cls="SP", target=meta[key].tolist(), gamma=5.0, mode="hard")
    add("sp_soft_g5_cg0_tensor_target", cls="SP", target=meta["composite"].clone(), gamma=5.0, mode="soft")
    add("sp_soft_g5_cg0_simclr_none", cls="SP", gamma=5.0, mode="soft")
    add("sp_soft_g5_cg0_simclr_range", cls="SP", target=list(range(n)), gamma=5.0, mode="soft")
    add("sp_soft_g5_cg1_trimask", cls="SP", mask=tri, gamma=5.0, mode="soft", correct_grad=True)
    add("sp_hard_g5_cg0_trimask", cls="SP", mask=tri, gamma=5.0, mode="hard")
    add("sp_default_gamma_partition", cls="SP", target=meta["partition"].tolist(), mode="hard")
    add("supcon1_partition", cls="SupConLoss1", target=meta["partition"].tolist())
    add("supcon1_simclr_none", cls="SupConLoss1")
    add("supcon1_trimask", cls="SupConLoss1", mask=tri)
    add("sp_soft_g5_cg0_partition_t0.2", cls="SP", target=meta["partition"].tolist(), gamma=5.0,
        mode="soft", temperature=0.2)
    store["cases"] = np.array(cases)
    np.savez_compressed(OUT / "cfg1_n64_d128.npz", **store)
    print("cfg1:", len(cases), "cases")

    # ---------------- ragged small case: n=5, d=16 (reference-scale batch tails) ----------------
    n, d = 5, 16
    g = torch.Generator().manual_seed(3)
    z1 = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1)
    z2 = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1)
    cases, store = [], dict(z1=z1.numpy(), z2=z2.numpy())
    labels = [0, 1, 0, 2, 1]
    store["labels"] = np.array(labels)
    for name, kw in {
        "sp_soft_g3": dict(cls="SP", target=labels, gamma=3.0, mode="soft"),
        "sp_hard_g3": dict(cls="SP", target=labels, gamma=3.0, mode="hard"),
        "supcon1": dict(cls="SupConLoss1", target=labels),
        "sp_simclr": dict(cls="SP", gamma=4.0, mode="soft"),
    }.items():
        res = run_reference(ref, z1, z2, **kw)
        for k, v in res.items():
            store[f"{name}/{k}"] = v
        cases.append(name)
    store["cases"] = np.array(cases)
    np.savez_compressed(OUT / "tiny_n5_d16.npz", **store)
    print("tiny:", len(cases), "cases")

    # ---------------- cfg2: n=256, d=256 ----------------
    n, d = 256, 256
    meta = wl.acdc_meta_labels(n)
    cases, store = [], {}
    for seed, key, gamma in ((0, "partition", 5.0), (1, "patient", 3.5), (2, "cycle", 2.0), (0, "composite", 5.0)):
        z1, z2 = wl.make_views(meta[key], d, sigma=0.7, seed=seed)
        name = f"sp_soft_g{gamma:g}_{key}_seed{seed}"
        store[f"{name}/z1"], store[f"{name}/z2"] = z1.numpy().astype(np.float32), z2.numpy().astype(np.float32)
        store[f"{name}/labels"] = meta[key].numpy()
        res = run_reference(ref, z1, z2, cls="SP", target=meta[key].tolist(), gamma=gamma, mode="soft")
        for k, v in res.items():
            store[f"{name}/{k}"] = v
        store[f"{name}/gamma"] = np.float64(gamma)
        cases.append(name)
    store["cases"] = np.array(cases)
    np.savez_compressed(OUT / "cfg2_n256_d256.npz", **store)
    print("cfg2:", len(cases), "cases")


def main_excl():
    """SupConLoss1(exclude_other_pos=True) (:97-100) -> tests/golden/excl_cases.npz (kept apart from the files
    above so that adding it does not rewrite them)."""
    ref = load_reference()
    wl = _load_workloads()
    store, cases = {}, []

    def add(name, z1, z2, **kw):
        store[f"{name}/z1"], store[f"{name}/z2"] = z1.numpy().astype(np.float32), z2.numpy().astype(np.float32)
        if kw.get("target") is not None:
            store[f"{name}/labels"] = np.asarray(kw["target"])
        if kw.get("mask") is not None:
            store[f"{name}/tri_mask"] = kw["mask"].numpy()
        store[f"{name}/temperature"] = np.float64(kw.get("temperature", 0.07))
        res = run_reference(ref, z1, z2, cls="SupConLoss1Excl", **kw)
        for k, v in res.items():
            store[f"{name}/{k}"] = v
        cases.append(name)

    n, d = 64, 128
    meta = wl.acdc_meta_labels(n)
    z1, z2 = wl.make_views(meta["partition"], d, sigma=0.7, seed=0)
    tri = torch.randint(0, 3, (n, n), generator=torch.Generator().manual_seed(7)).float()
    tri.fill_diagonal_(1.0)
    for key in ("partition", "patient", "cycle", "composite"):
        add(f"n64_{key}", z1, z2, target=meta[key].tolist())
    add("n64_simclr_none", z1, z2)
    add("n64_trimask", z1, z2, mask=tri)
    add("n64_partition_t0.2", z1, z2, target=meta["partition"].tolist(), temperature=0.2)
    g = torch.Generator().manual_seed(3)
    t1 = torch.nn.functional.normalize(torch.randn(5, 16, generator=g), dim=1)
    t2 = torch.nn.functional.normalize(torch.randn(5, 16, generator=g), dim=1)
    add("n5_ragged", t1, t2, target=[0, 1, 0, 2, 1])
    n, d = 256, 256
    meta = wl.acdc_meta_labels(n)
    z1, z2 = wl.make_views(meta["patient"], d, sigma=0.7, seed=1)
    add("n256_patient", z1, z2, target=meta["patient"].tolist())
    store["cases"] = np.array(cases)
    np.savez_compressed(OUT / "excl_cases.npz", **store)
    print("excl:", len(cases), "cases")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "excl":
        main_excl()
    else:
        main()
        main_excl()
