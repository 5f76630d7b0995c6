"""Development tool: where the host time of a small-batch loss call goes (cProfile over eager cfg2 steps).

    python tools/gpu_small_profile.py [grouped]
"""
import cProfile
import pathlib
import pstats
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import spcl_b200  # noqa: E402
from spcl_b200.workloads import acdc_meta_labels, make_views  # noqa: E402

n, d = 256, 256
meta = acdc_meta_labels(n)
probs = []
for k, g in zip(("partition", "patient", "cycle"), (5.0, 3.5, 2.0)):
    z1, z2 = make_views(meta[k], d, sigma=0.7, seed=1)
    crit = spcl_b200.SelfPacedSupConLoss(weight_update="soft", correct_grad=True, check_nan=False, validate=False)
    crit.set_gamma(g)
    probs.append((crit, z1.cuda().requires_grad_(True), z2.cuda().requires_grad_(True), meta[k].int().cuda()))
grouped = len(sys.argv) > 1


def step():
    if grouped:
        for _, a, b, _l in probs:
            a.grad = b.grad = None
        ls = spcl_b200.grouped_forward([q[0] for q in probs], [(q[1], q[2]) for q in probs], [q[3] for q in probs])
        (ls[0] + ls[1] + ls[2]).backward()
        return
    total = None
    for crit, a, b, lab in probs:
        a.grad = b.grad = None
        l = crit(a, b, target=lab)
        total = l if total is None else total + l
    total.backward()


for _ in range(20):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
