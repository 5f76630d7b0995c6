"""Decoder pre-training step: dense pixel contrast on a decoder feature map of the reference's UNet.

    python tools/decoder_step.py [--batch 16] [--steps 10] > gpurun_out/decoder_step.json

What the reference does per batch in this stage (main_pretrain_decoder.py:42-76, semi_seg/hooks/infonce.py:198-241,
config/hooks/infonce_dense.yaml): the UNet runs up to ``Up_conv3`` with everything up to ``Conv5`` frozen
(``model.set_grad(False)`` / ``set_grad(True, start="Conv5", end=until, include_start=False)``), the feature map of
both views goes through ``DenseProjectionHead(input_dim, 256, 256, "mlp", normalize=True, spatial_size=(10, 10))``,
``region_extractor`` picks 5 pooled pixels per image (same coordinates for both views: ``FixRandomSeed(seed)``), and
``SupConLoss1`` contrasts them with ``target=list(range(N))``; backward; optimiser step.  The affine transform between
the views is the data pipeline's and is left out (identity).

Arms (identical weights, inputs and coordinates):
  reference : baseline/_ref UNet + DenseProjectionHead + the restated region_extractor + contrast_loss3.SupConLoss1,
              ``loss.item()`` per batch
  fused     : the same UNet, this repo's DenseProjectionHead (1x1 convs, then pool + normalise + point gather as ONE
              kernel that only reads the sampled windows) + fused SupConLoss1
"""
import argparse
import json
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import spcl_b200                                            # noqa: E402
from spcl_b200 import hooks                                 # noqa: E402
from spcl_b200.dense import point_coordinates               # noqa: E402


def _reference_parts():
    try:
        from baseline import ref_loader
        if ref_loader.available():
            return ref_loader.unet_module().UNet, ref_loader.heads_module().DenseProjectionHead, ref_loader.loss_module()
    except Exception as e:                                   # noqa: BLE001
        print(f"decoder_step: reference files unavailable ({type(e).__name__}: {e})", file=sys.stderr)
    return None, None, None


def _region_extractor_ref(features, pts, w):
    """infonce.py:233-241 with the coordinates drawn beforehand (so both arms use the same ones)."""
    out = []
    for fmap, row in zip(features, pts.tolist()):
        out.append(torch.stack([fmap[:, q // w, q % w] for q in row], dim=0))
    return torch.cat(out, dim=0)


def measure(batch=16, steps=10, warmup=3, image=224, until="Up_conv3", spatial=(10, 10), point_nums=5,
            pool_name="adaptive_avg", tf32=True, device="cuda"):
    unet_cls, ref_head_cls, ref_loss = _reference_parts()
    if unet_cls is None:
        return {"unavailable": "baseline/_ref is not installed (tools/install_ref.sh)"}
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    out = {"workload": f"decoder_step_2x{batch}_{image}x{image}_{until}_{spatial[0]}x{spatial[1]}_{point_nums}pts",
           "steps": steps, "pool_name": pool_name, "anchors_N": 2 * batch * point_nums}
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(2 * batch, 1, image, image, generator=gen).to(device)
    for arm in ("reference", "fused"):
        torch.manual_seed(0)
        net = unet_cls(input_dim=1, num_classes=4, max_channel=256).to(device)
        cdim = net.get_channel_dim(until)
        kw = dict(input_dim=cdim, hidden_dim=256, output_dim=256, head_type="mlp", normalize=True, pool_name=pool_name,
                  spatial_size=spatial)
        head = (ref_head_cls(**kw) if arm == "reference" else hooks.DenseProjectionHead(**kw)).to(device)
        crit = ref_loss.SupConLoss1() if arm == "reference" else spcl_b200.SupConLoss1(check_nan=False, validate=False)
        net.requires_grad_(False)                                    # set_grad(False) ...
        trainable = []
        for name in ("Up5", "Up_conv5", "Up4", "Up_conv4", "Up3", "Up_conv3", "Up2", "Up_conv2"):
            getattr(net, "_" + name).requires_grad_(True)            # ... set_grad(True, start="Conv5", end=until, include_start=False)
            trainable += list(getattr(net, "_" + name).parameters())
            if name == until:
                break
        params = trainable + list(head.parameters())
        opt = torch.optim.SGD(params, lr=1e-3, momentum=0.9)
        losses, gnorms = [], []

        def step(k):
            opt.zero_grad(set_to_none=True)
            feat = net(x, until=until)
            pts = point_coordinates(batch, spatial[0], spatial[1], point_nums, seed=100 + k)
            if arm == "reference":
                za, zb = torch.chunk(head(feat), 2)
                a = _region_extractor_ref(za, pts, spatial[1])
                b = _region_extractor_ref(zb, pts, spatial[1])
                loss = crit(a, b, target=list(range(a.shape[0])))
                losses.append(loss.item())                          # infonce.py:219
            else:
                rows = head.rows(feat, points=torch.cat([pts, pts]))
                a, b = torch.chunk(rows, 2)
                loss = crit(a, b)                                   # target=range(n) == identity mask (:140-143)
                losses.append(loss.detach())
            loss.backward()
            gnorms.append(torch.sqrt(sum((q.grad.double() ** 2).sum() for q in params if q.grad is not None)).detach())
            opt.step()

        for k in range(warmup):
            step(k)
        torch.cuda.synchronize()
        losses.clear(); gnorms.clear()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for k in range(steps):
            step(warmup + k)
        t1.record()
        torch.cuda.synchronize()
        out[arm] = {"ms_per_step": t0.elapsed_time(t1) / steps, "losses": [float(v) for v in losses[:5]],
                    "grad_norms": [float(v) for v in gnorms[:5]]}
        del net, head, opt
        torch.cuda.empty_cache()
    out["step_speedup_fused_vs_reference"] = out["reference"]["ms_per_step"] / out["fused"]["ms_per_step"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--image", type=int, default=224)
    ap.add_argument("--pool", default="adaptive_avg")
    args = ap.parse_args()
    print(json.dumps(measure(args.batch, args.steps, args.warmup, args.image, pool_name=args.pool)))


if __name__ == "__main__":
    main()
