"""cfg5 (BASELINE.json configs[4]): encoder pre-training step on synthetic 224x224 ACDC-shape slices, batch 64
(2 views -> 128 images), fused loss drop-in vs the reference-style dense loss on the same backbone.

    python tools/cfg5_step.py [--batch 64] [--steps 10] > gpurun_out/cfg5.json

The backbone is spcl_b200.workloads.acdc_encoder (stand-in for UNet(..., until="Conv5"), semi_seg/arch/unet.py),
the projector ProjectionHead(256, 256, 256, "mlp") (infonce.py:96-99); a step = forward, loss, backward, Adam step
(new_pretrain.py:52-96 + infonce.py:171-195).  The "reference" arm runs the oracle's dense fp32 port of
contrast_loss3.py on the GPU (checker code, timed here only as the thing the drop-in replaces) and reads
loss.item() every step like the reference hook does (infonce.py:183).
"""
import argparse
import json
import pathlib
import sys

import torch
import torch.nn.functional as F

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import spcl_b200                                            # noqa: E402
from spcl_b200 import hooks                                 # noqa: E402
from spcl_b200.workloads import acdc_encoder, acdc_meta_labels   # noqa: E402
from oracle.dense_port import dense_supcon                  # noqa: E402  (baseline arm only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    n = args.batch
    labels = acdc_meta_labels(n)["partition"].tolist()
    out = {"workload": f"cfg5_encoder_step_2x{n}_224x224_d256", "steps": args.steps}
    x = torch.randn(2 * n, 1, 224, 224, device="cuda")
    for arm in ("fused", "fused_graph", "reference_style"):
        torch.manual_seed(0)
        enc = acdc_encoder(1, 256).cuda()
        head = hooks.ProjectionHead(input_dim=256, hidden_dim=256, output_dim=256, head_type="mlp", normalize=True).cuda()
        params = list(enc.parameters()) + list(head.parameters())
        opt = torch.optim.Adam(params, lr=1e-6)
        crit = spcl_b200.SelfPacedSupConLoss(weight_update="soft", correct_grad=True, check_nan=False, validate=False,
                                             cuda_graph=(arm == "fused_graph"))
        crit.set_gamma(8.0)
        meter = hooks.DeviceMeter()
        losses, loss_ms = [], []

        def step(timed):
            opt.zero_grad(set_to_none=True)
            feat = enc(x)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if arm != "reference_style":
                za, zb = torch.chunk(head(feat), 2)
                loss = crit(za, zb, target=labels)
                meter.add(loss)
            else:
                z = head._header[:-1](feat)
                za, zb = torch.chunk(F.normalize(z, p=2, dim=1), 2)
                loss = dense_supcon(za, zb, target=labels, gamma=8.0, mode="soft", correct_grad=True).loss
                losses.append(loss.item())                      # the reference hook's per-batch sync
            e1.record()
            loss.backward()
            opt.step()
            return e0, e1

        for _ in range(args.warmup):
            step(False)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        evs = [step(True) for _ in range(args.steps)]
        t1.record()
        torch.cuda.synchronize()
        out[arm] = {"ms_per_step": t0.elapsed_time(t1) / args.steps,
                    "head_plus_loss_fwd_ms": sum(a.elapsed_time(b) for a, b in evs) / args.steps,
                    "loss": meter.summary() if arm != "reference_style" else sum(losses[-args.steps:]) / args.steps}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
