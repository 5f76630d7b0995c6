"""cfg5 (BASELINE.json configs[4]): encoder pre-training step on synthetic 224x224 ACDC-shape slices, batch 64
(2 views -> 128 images), fused loss drop-in vs the UNMODIFIED reference loss on the same backbone.

    python tools/cfg5_step.py [--batch 64] [--steps 10] > gpurun_out/cfg5.json

Backbone: the reference's own ``UNet(input_dim=1, num_classes=4, max_channel=256)(x, until="Conv5")``
(semi_seg/arch/unet.py:156-230) and projector ``ProjectionHead(256, 256, 256, "mlp", normalize=True)``
(contrastyou/projectors/heads.py:76-92), both loaded from baseline/_ref/ (tools/install_ref.sh; byte-identical copies,
checked against baseline/REF_MANIFEST.sha256).  A step = forward, loss, backward, Adam step (the ~40 lines of
main_pretrain_encoder.py:41-74 / new_pretrain.py:52-96 + infonce.py:171-195 that are on the path).

Arms (identical initial weights, identical inputs):
  reference : reference UNet + reference ProjectionHead + reference SelfPacedSupConLoss (contrast_loss3.py) and the
              reference hook's per-batch ``loss.item()`` (infonce.py:183)
  fused     : reference UNet + this repo's ProjectionHead tail + fused loss, meters kept on the device
  fused_graph : the same with ``cuda_graph=True`` (loss fwd + bwd as one graph replay)
If baseline/_ref is absent the backbone falls back to the from-scratch stand-in (workloads.acdc_encoder) and the
reference arm to the oracle's dense port; the output says which.
"""
import argparse
import json
import pathlib
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import spcl_b200                                            # noqa: E402
from spcl_b200 import hooks                                 # noqa: E402
from spcl_b200.workloads import acdc_encoder, acdc_meta_labels   # noqa: E402


def _reference_parts():
    try:
        from baseline import ref_loader
        if ref_loader.available():
            return ref_loader.unet_module().UNet, ref_loader.heads_module().ProjectionHead, ref_loader.loss_module()
    except Exception as e:                                   # noqa: BLE001
        print(f"cfg5: reference files unavailable ({type(e).__name__}: {e})", file=sys.stderr)
    return None, None, None


class _Encoder(torch.nn.Module):
    def __init__(self, unet_cls):
        super().__init__()
        self.kind = "reference UNet (baseline/_ref/semi_seg/arch/unet.py)" if unet_cls else "stand-in acdc_encoder"
        self.net = unet_cls(input_dim=1, num_classes=4, max_channel=256) if unet_cls else acdc_encoder(1, 256)
        self._ref = unet_cls is not None

    def forward(self, x):
        return self.net(x, until="Conv5") if self._ref else self.net(x)


def measure(batch: int = 64, steps: int = 10, warmup: int = 3, arms=("reference", "fused", "fused_graph"),
            image: int = 224, device="cuda"):
    unet_cls, ref_head_cls, ref_loss = _reference_parts()
    n = batch
    labels = acdc_meta_labels(n)["partition"].tolist()
    out = {"workload": f"cfg5_encoder_step_2x{n}_{image}x{image}_d256", "steps": steps, "optimizer": "Adam lr=1e-6",
           "hyper": {"tau": 0.07, "gamma": 8.0, "mode": "soft", "correct_grad": True, "labels": "partition"}}
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(2 * n, 1, image, image, generator=gen).to(device)
    for arm in arms:
        torch.manual_seed(0)
        enc = _Encoder(unet_cls).to(device)
        out["backbone"] = enc.kind
        if arm == "reference" and ref_head_cls is not None:
            head = ref_head_cls(input_dim=256, hidden_dim=256, output_dim=256, head_type="mlp", normalize=True).to(device)
        else:
            head = hooks.ProjectionHead(input_dim=256, hidden_dim=256, output_dim=256, head_type="mlp",
                                        normalize=True).to(device)
        params = list(enc.parameters()) + list(head.parameters())
        opt = torch.optim.Adam(params, lr=1e-6)
        if arm == "reference":
            if ref_loss is not None:
                crit = ref_loss.SelfPacedSupConLoss(temperature=0.07, weight_update="soft", correct_grad=True)
                out["reference_loss"] = "unmodified contrast_loss3.SelfPacedSupConLoss (baseline/_ref)"
            else:
                from oracle.dense_port import dense_supcon          # checker code standing in for the reference file

                class _Port(torch.nn.Module):
                    gamma = 8.0

                    def set_gamma(self, g):
                        self.gamma = g

                    def forward(self, a, b, target=None):
                        return dense_supcon(a, b, target=target, gamma=self.gamma, mode="soft", correct_grad=True).loss
                crit = _Port()
                out["reference_loss"] = "oracle dense port (baseline/_ref absent)"
        else:
            crit = spcl_b200.SelfPacedSupConLoss(weight_update="soft", correct_grad=True, check_nan=False,
                                                 validate=False, cuda_graph=(arm == "fused_graph"))
        crit.set_gamma(8.0)
        meter = hooks.DeviceMeter()
        losses, gnorms = [], []

        def step():
            opt.zero_grad(set_to_none=True)
            feat = enc(x)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            za, zb = torch.chunk(head(feat), 2)
            loss = crit(za, zb, target=labels)
            if arm == "reference":
                losses.append(loss.item())                      # the reference hook's per-batch sync (infonce.py:183)
            else:
                meter.add(loss)
                losses.append(loss.detach())
            e1.record()
            loss.backward()
            gnorms.append(torch.linalg.vector_norm(torch.stack([p.grad.norm() for p in head.parameters()])).detach())
            opt.step()
            return e0, e1

        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        losses.clear(); gnorms.clear()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        evs = [step() for _ in range(steps)]
        t1.record()
        torch.cuda.synchronize()
        out[arm] = {"ms_per_step": t0.elapsed_time(t1) / steps,
                    "head_plus_loss_fwd_ms": sum(a.elapsed_time(b) for a, b in evs) / steps,
                    "losses": [float(v) for v in losses[:5]],
                    "head_grad_norms": [float(v) for v in gnorms[:5]]}
        del enc, head, opt, crit
        torch.cuda.empty_cache()
    if "reference" in out and "fused" in out:
        a, b = out["reference"], out["fused"]
        out["max_loss_rel_diff_first_steps"] = max(abs(p - q) / abs(p) for p, q in zip(a["losses"], b["losses"]))
        out["max_grad_norm_rel_diff_first_steps"] = max(abs(p - q) / abs(p) for p, q in
                                                        zip(a["head_grad_norms"], b["head_grad_norms"]))
        out["step_speedup_fused_vs_reference"] = a["ms_per_step"] / b["ms_per_step"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--image", type=int, default=224)
    args = ap.parse_args()
    print(json.dumps(measure(args.batch, args.steps, args.warmup, image=args.image)))


if __name__ == "__main__":
    main()
