"""Synthetic anchors and meta-labels of the shapes BASELINE.json names.

Host-side helpers shared by the tests and ``bench.py``.  Labels follow the
reference's batch composition: ``ContrastBatchSampler`` draws scans x 3 partitions
(``semi_seg/data/rearr.py:59-75``, ``contrastyou/data/dataset.py:34-43``) and the
label generators map them to integer lists (``semi_seg/epochers/helper.py:48-65``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

__all__ = ["acdc_meta_labels", "clustered_embeddings", "make_views", "WORKLOADS", "make_workload", "acdc_encoder"]


def acdc_meta_labels(n: int) -> dict:
    """anchor i -> patient = i // 6, phase = (i // 3) % 2, partition = i % 3."""
    i = torch.arange(n, dtype=torch.int64)
    patient, phase, partition = i // 6, (i // 3) % 2, i % 3
    return {
        "partition": partition,
        "patient": patient,
        "cycle": phase,
        "self": i.clone(),
        # one packed key: equal iff all three fields are equal (< 2**24 for any sane n)
        "composite": partition + 3 * (phase + 2 * patient),
    }


def clustered_embeddings(labels: torch.Tensor, d: int, sigma: float, gen: torch.Generator,
                         dtype=torch.float32) -> torch.Tensor:
    """normalize(centroid[label] + sigma * randn): spreads l_ij over the gamma range."""
    labels = labels.to(torch.int64)
    k = int(labels.max().item()) + 1
    centroids = torch.randn(k, d, generator=gen, dtype=dtype)
    x = centroids[labels] + sigma * torch.randn(labels.numel(), d, generator=gen, dtype=dtype)
    return F.normalize(x, dim=1)


def make_views(labels: torch.Tensor, d: int, *, sigma: float = 0.7, seed: int = 0):
    """Two views (z1, z2), each [n, d] fp32 unit rows, sharing the label centroids."""
    gen = torch.Generator().manual_seed(seed)
    labels = labels.to(torch.int64)
    k = int(labels.max().item()) + 1
    centroids = torch.randn(k, d, generator=gen)
    z = []
    for _ in range(2):
        x = centroids[labels] + sigma * torch.randn(labels.numel(), d, generator=gen)
        z.append(F.normalize(x, dim=1))
    return z[0], z[1]


# name -> (n per view, d, label kind)
WORKLOADS = {
    # cfg1: the reference's own CPU-runnable case
    "cfg1_cpu_2x64_d128": dict(n=64, d=128, labels="partition"),
    # cfg2: encoder global contrast, Conv5 -> ProjectionHead width (infonce.py:97)
    "cfg2_encoder_2x256_d256": dict(n=256, d=256, labels="composite"),
    # cfg3: dense decoder pixel contrast, 16 slices x 32x32 pixels per view
    "cfg3_dense_2x16384_d128_simclr": dict(n=16384, d=128, labels="self"),
    "cfg3_dense_2x16384_d128_slice": dict(n=16384, d=128, labels="slice1024"),
    # cfg4: row-sharded dense contrast
    "cfg4_dense_2x131072_d128_simclr": dict(n=131072, d=128, labels="self"),
}


def make_workload(name: str, *, seed: int = 0, sigma: float = 0.7):
    """-> (z1, z2, labels[int64, n]) on CPU."""
    spec = WORKLOADS[name]
    n, d, kind = spec["n"], spec["d"], spec["labels"]
    if kind == "slice1024":
        labels = torch.arange(n, dtype=torch.int64) // 1024
    else:
        labels = acdc_meta_labels(n)[kind]
    if kind == "self":
        # one centroid per anchor would be n x d randn twice; same statistics, cheaper:
        gen = torch.Generator().manual_seed(seed)
        base = torch.randn(n, d, generator=gen)
        z1 = F.normalize(base + sigma * torch.randn(n, d, generator=gen), dim=1)
        z2 = F.normalize(base + sigma * torch.randn(n, d, generator=gen), dim=1)
    else:
        z1, z2 = make_views(labels, d, sigma=sigma, seed=seed)
    return z1, z2, labels


def acdc_encoder(input_dim: int = 1, max_channel: int = 256, momentum: float = 0.1):
    """Stand-in for the cfg5 backbone, ``UNet(input_dim=1, num_classes=4, max_channel=256)(x, until="Conv5")``
    (``semi_seg/arch/unet.py:100-170``): five double 3x3 conv + BatchNorm + ReLU stages of widths
    max_channel/16 * (1, 2, 4, 8, 16) with 2x2 max-pools between them.  Written from the architecture's
    description for the benchmark harness only (the UNet itself is outside the hot path);
    plain torch / cuDNN layers, random init."""
    from torch import nn

    def stage(cin, cout):
        return nn.Sequential(
            nn.Conv2d(cin, cout, 3, 1, 1, bias=False), nn.BatchNorm2d(cout, momentum=momentum), nn.ReLU(inplace=True),
            nn.Conv2d(cout, cout, 3, 1, 1, bias=False), nn.BatchNorm2d(cout, momentum=momentum), nn.ReLU(inplace=True))

    widths = [max_channel // 16 * m for m in (1, 2, 4, 8, 16)]
    layers, cin = [], input_dim
    for k, cout in enumerate(widths):
        if k:
            layers.append(nn.MaxPool2d(2, 2))
        layers.append(stage(cin, cout))
        cin = cout
    return nn.Sequential(*layers)
