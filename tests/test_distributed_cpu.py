"""Row-sharding plumbing on CPU: world_size-2 gloo, with the oracle standing in for the kernels.

Checks the parts of spcl_b200.distributed that do not depend on CUDA: the global anchor order, the
all-gathers, the all-reduce of the three partial sums, the row ranges handed to the kernels and the
gradient rows handed back -- against the single-process oracle on the concatenated batch.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.closed_form import supcon_closed_form


class OracleBackend:
    """Stands in for NativeBackend in CPU tests only."""
    name = "oracle"

    def check(self, plan, d):
        pass

    def pack(self, z1, z2):
        return torch.cat([z1, z2]).double()

    def _full(self, z_all, labels_all, inv_tau, gamma, mode, **kw):
        N = z_all.shape[0]
        z = z_all.numpy()
        return supcon_closed_form(z[: N // 2], z[N // 2:], anchor_labels=labels_all.numpy(), temperature=1.0 / inv_tau,
                                  gamma=gamma, mode={0: "none", 1: "hard", 2: "soft"}[mode], **kw)

    def forward_rows(self, z_all, labels_all, plan, inv_tau, gamma, mode, group=None):
        r = self._full(z_all, labels_all, inv_tau, gamma, mode, want_grad=False,
                       row_range=(plan.row_begin, plan.row_end))
        sl = slice(plan.row_begin, plan.row_end)
        A = r["wp"][sl] / r["c"][sl]
        stats = np.stack([r["logD"][sl], 1.0 / r["c"][sl], A, A * np.exp(inv_tau - r["logD"][sl])], axis=0)
        p = r["partial"]
        return (torch.from_numpy(stats), torch.tensor([p["loss_sum"], p["wp_sum"], p["p_sum"]], dtype=torch.float64),
                torch.zeros(1))

    def finalize(self, partials, N, correct_grad):
        loss_sum, wp, pc = partials.tolist()
        ratio = wp / pc
        scale = 1.0 / ratio if (correct_grad and ratio > 0) else 1.0
        return torch.tensor([-(loss_sum / N) * scale, ratio, scale, scale / N], dtype=torch.float64)

    def backward_rows(self, z_all, labels_all, sig, stats_all, scalars, grad, plan, inv_tau, gamma, mode, d):
        # the oracle recomputes everything; scale comes from the already reduced scalars
        r = self._full(z_all, labels_all, inv_tau, gamma, mode)
        dz = np.concatenate([r["dz1"], r["dz2"]])[plan.row_begin:plan.row_end]
        return torch.from_numpy(dz) * float(scalars[2]) * float(grad[0])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_loc, d, mode, gamma, correct_grad, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from spcl_b200.distributed import sharded_supcon_loss
        g = torch.Generator().manual_seed(7)
        n = n_loc * world
        labels = torch.randint(0, 5, (n,), generator=g).int()
        z1 = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1).double()
        z2 = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1).double()
        sl = slice(rank * n_loc, (rank + 1) * n_loc)
        a = z1[sl].clone().requires_grad_(True)
        b = z2[sl].clone().requires_grad_(True)
        loss, scalars = sharded_supcon_loss(a, b, labels[sl].contiguous(), temperature=0.1, gamma=gamma, mode=mode,
                                            correct_grad=correct_grad, backend=OracleBackend())
        (loss * 2.0).backward()
        ref = supcon_closed_form(z1.numpy(), z2.numpy(), target=labels.numpy(), temperature=0.1, gamma=gamma,
                                 mode={0: "none", 1: "hard", 2: "soft"}[mode], correct_grad=correct_grad, grad_out=2.0)
        ok = (np.isclose(loss.item(), ref["loss"], rtol=1e-10)
              and np.isclose(scalars[1].item(), ref["ratio"], rtol=1e-10)
              and np.allclose(a.grad.numpy(), ref["dz1"][sl], rtol=1e-8, atol=1e-14)
              and np.allclose(b.grad.numpy(), ref["dz2"][sl], rtol=1e-8, atol=1e-14))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode,gamma,cg", [(0, 1e6, False), (2, 3.0, True), (1, 2.5, False)])
def test_row_sharded_matches_single_process(mode, gamma, cg):
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), 12, 16, mode, gamma, cg, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_shard_plan_ranges():
    from spcl_b200.distributed import ShardPlan
    plans = [ShardPlan(64, 4, r) for r in range(4)]
    assert [p.row_begin for p in plans] == [0, 128, 256, 384] and plans[-1].row_end == plans[0].N == 512


def test_native_backend_rejects_unaligned_shards():
    from spcl_b200.distributed import NativeBackend, ShardPlan
    from spcl_b200 import SpclError
    with pytest.raises(SpclError):
        NativeBackend().check(ShardPlan(50, 2, 0), 128)
