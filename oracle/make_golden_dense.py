"""Generate tests/golden/dense_points.npz from the reference's own dense-hook sampling code.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python oracle/make_golden_dense.py

``semi_seg/hooks/infonce.py`` cannot be imported (its ``deepclustering2`` / ``matplotlib.pyplot`` imports are
absent), so the two functions on this path -- ``get_n_point_coordinate`` (:20-22) and the static method
``_INFONCEDenseHook.region_extractor`` (:233-241) -- are cut out of the UNMODIFIED source file with ``ast`` and
executed as they are.  ``FixRandomSeed(seed)`` (``deepclustering2``, un-pinned and absent) is stood in for by
``numpy.random.seed(seed)`` before each call, which is the part of it the draw depends on.  The feature maps go
through the DenseProjectionHead tail exactly as heads.py:112-114 spells it (``AdaptiveAvgPool2d`` then
``F.normalize(dim=1)``; both are torch, not reference code).
"""
from __future__ import annotations

import ast
import pathlib

import numpy as np
import torch
import torch.nn.functional as F

ROOT = pathlib.Path(__file__).resolve().parent.parent
REF_FILE = pathlib.Path("/root/reference/semi_seg/hooks/infonce.py")
OUT = ROOT / "tests" / "golden" / "dense_points.npz"


def load_reference_functions():
    src = REF_FILE.read_text()
    tree = ast.parse(src)
    ns = {"np": np, "torch": torch}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "get_n_point_coordinate":
            exec(compile(ast.Module([node], []), str(REF_FILE), "exec"), ns)
        if isinstance(node, ast.ClassDef) and node.name == "_INFONCEDenseHook":
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name == "region_extractor":
                    item.decorator_list = []            # drop @staticmethod: called as a plain function
                    exec(compile(ast.Module([item], []), str(REF_FILE), "exec"), ns)
    return ns["get_n_point_coordinate"], ns["region_extractor"]


# name -> (B, C, H, W, pooled (ph, pw), point_nums, seed)
CASES = {
    "nopool_12": (3, 8, 12, 12, (12, 12), 5, 0),
    "pool_28_to_10": (4, 16, 28, 28, (10, 10), 5, 7),          # the decoder default spatial_size (infonce.py:76)
    "pool_32_to_16": (2, 32, 32, 32, (16, 16), 5, 123),
    "rect_20x14_to_7x5": (2, 5, 20, 14, (7, 5), 3, 42),
}


def main():
    get_pts, region_extractor = load_reference_functions()
    out = {"cases": np.array(list(CASES))}
    for name, (B, C, H, W, (ph, pw), P, seed) in CASES.items():
        gen = torch.Generator().manual_seed(seed)
        x = torch.randn(B, C, H, W, generator=gen)
        normed = F.normalize(torch.nn.AdaptiveAvgPool2d((ph, pw))(x), p=2, dim=1)
        np.random.seed(seed)
        rows = region_extractor(normed, point_nums=P)
        np.random.seed(seed)
        pts = [[a * pw + b for a, b in get_pts(h=ph, w=pw, n=P)] for _ in range(B)]
        out[f"{name}/x"] = x.numpy()
        out[f"{name}/geom"] = np.array([ph, pw, P, seed], dtype=np.int64)
        out[f"{name}/points"] = np.array(pts, dtype=np.int32)
        out[f"{name}/rows"] = rows.numpy()
        out[f"{name}/all_rows"] = normed.permute(0, 2, 3, 1).reshape(-1, C).numpy()
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT} ({OUT.stat().st_size} bytes)")


if __name__ == "__main__":
    main()
