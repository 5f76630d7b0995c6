import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """Lazy view of one tests/golden/*.npz (made by oracle/make_golden.py from the real reference)."""

    def __init__(self, name):
        self._z = np.load(GOLDEN / name, allow_pickle=False)
        self.cases = [str(c) for c in self._z["cases"]]

    def __getitem__(self, key):
        return self._z[key]

    def __contains__(self, key):
        return key in self._z.files

    def case(self, name):
        return {k: self._z[f"{name}/{k}"] for k in ("loss", "ratio", "dz1", "dz2")}


@pytest.fixture(scope="session")
def golden_cfg1():
    return Golden("cfg1_n64_d128.npz")


@pytest.fixture(scope="session")
def golden_cfg2():
    return Golden("cfg2_n256_d256.npz")


@pytest.fixture(scope="session")
def golden_tiny():
    return Golden("tiny_n5_d16.npz")


def parse_cfg1_case(name, g):
    """case name -> kwargs for the oracle / product (mirrors oracle/make_golden.py)."""
    kw = dict(temperature=0.07, gamma=1e6, mode="hard", correct_grad=False, target=None, mask=None, cls="SP")
    if name.startswith("supcon1"):
        kw.update(cls="SupConLoss1", mode="none")
    parts = name.split("_")
    for p in parts:
        if p in ("hard", "soft"):
            kw["mode"] = p
        elif p.startswith("g") and p[1:].replace(".", "").replace("e+", "").isdigit():
            kw["gamma"] = float(p[1:])
        elif p.startswith("cg") and p[2:].isdigit():
            kw["correct_grad"] = bool(int(p[2:]))
        elif p.startswith("t0."):
            kw["temperature"] = float(p[1:])
    if "trimask" in name:
        kw["mask"] = g["tri_mask"]
    elif "simclr_none" in name:
        pass
    elif "simclr_range" in name:
        kw["target"] = list(range(g["z1"].shape[0]))
    elif "tensor_target" in name:
        kw["target"] = g["labels_composite"]
    else:
        for key in ("partition", "patient", "cycle", "composite"):
            if key in parts:
                kw["target"] = g["labels_" + key].tolist()
    return kw


def excl_case_inputs(g, name):
    """inputs of one ``excl_cases.npz`` case (SupConLoss1(exclude_other_pos=True), oracle/make_golden.py)."""
    kw = dict(temperature=float(g[f"{name}/temperature"]), target=None, mask=None)
    if f"{name}/tri_mask" in g:
        kw["mask"] = g[f"{name}/tri_mask"]
    elif f"{name}/labels" in g:
        kw["target"] = g[f"{name}/labels"].tolist()
    return g[f"{name}/z1"], g[f"{name}/z2"], kw
