"""Row-sharded loss on real GPUs (NCCL): R-rank result == 1-rank result on the concatenated batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_loc, d, mode, gamma, cg, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import spcl_b200
        from spcl_b200.distributed import sharded_supcon_loss
        g = torch.Generator().manual_seed(11)
        n = n_loc * world
        labels = torch.randint(0, 37, (n,), generator=g).int()
        cent = torch.randn(37, d, generator=g)
        z1 = torch.nn.functional.normalize(cent[labels.long()] + 0.7 * torch.randn(n, d, generator=g), dim=1)
        z2 = torch.nn.functional.normalize(cent[labels.long()] + 0.7 * torch.randn(n, d, generator=g), dim=1)
        sl = slice(rank * n_loc, (rank + 1) * n_loc)
        a = z1[sl].cuda().requires_grad_(True)
        b = z2[sl].cuda().requires_grad_(True)
        loss, scalars = sharded_supcon_loss(a, b, labels[sl].cuda(), temperature=0.07, gamma=gamma, mode=mode,
                                            correct_grad=cg)
        loss.backward()
        # single-GPU run of the same kernels on the concatenated batch (rank-local, no collectives)
        A = z1.cuda().requires_grad_(True)
        B = z2.cuda().requires_grad_(True)
        ref, sc, _ = spcl_b200.supcon_loss(A, B, target=labels.cuda(), temperature=0.07, gamma=gamma, mode=mode,
                                           correct_grad=cg, precision="bf16")
        ref.backward()
        ok = (np.isclose(loss.item(), ref.item(), rtol=2e-5)
              and np.isclose(scalars[1].item(), sc[1].item(), rtol=2e-5)
              and torch.allclose(a.grad, A.grad[sl], rtol=2e-3, atol=2e-3 * A.grad.abs().max().item())
              and torch.allclose(b.grad, B.grad[sl], rtol=2e-3, atol=2e-3 * B.grad.abs().max().item()))
        out[rank] = (bool(ok), loss.item(), ref.item())
    finally:
        dist.destroy_process_group()


def _oracle_worker(rank, world, port, n_loc, d, mode, gamma, cg, out):
    """Sharded result vs the fp64 numpy ORACLE (not vs the product): every rank's loss / ratio / gradient rows against
    ``supcon_closed_form`` on the bf16-rounded operands of the concatenated batch.  Shards straddle many row tiles
    (2 * n_loc / 128 per rank) and the slice-style labels put positives on other ranks."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oracle.closed_form import supcon_closed_form
        from spcl_b200.distributed import sharded_supcon_loss
        g = torch.Generator().manual_seed(23)
        n = n_loc * world
        n_cls = 61
        labels = ((torch.arange(n) // 48) % n_cls).int()            # runs of 48 samples, classes recur on every rank
        cent = torch.randn(n_cls, d, generator=g)
        z1 = torch.nn.functional.normalize(cent[labels.long()] + 0.7 * torch.randn(n, d, generator=g), dim=1)
        z2 = torch.nn.functional.normalize(cent[labels.long()] + 0.7 * torch.randn(n, d, generator=g), dim=1)
        z1, z2 = z1.bfloat16().float(), z2.bfloat16().float()       # the operands the tensor-core kernels consume
        sl = slice(rank * n_loc, (rank + 1) * n_loc)
        a = z1[sl].cuda().requires_grad_(True)
        b = z2[sl].cuda().requires_grad_(True)
        loss, scalars = sharded_supcon_loss(a, b, labels[sl].cuda(), temperature=0.07, gamma=gamma, mode=mode,
                                            correct_grad=cg)
        loss.backward()
        ref = supcon_closed_form(z1.double().numpy(), z2.double().numpy(), target=labels.numpy(), gamma=gamma,
                                 mode={0: "none", 1: "hard", 2: "soft"}[mode], correct_grad=cg)
        got = np.concatenate([a.grad.cpu().numpy(), b.grad.cpu().numpy()]).astype(np.float64)
        want = np.concatenate([ref["dz1"][sl], ref["dz2"][sl]])
        rel = np.abs(got - want).max() / np.abs(want).max()
        cos = (got * want).sum() / (np.linalg.norm(got) * np.linalg.norm(want))
        ok = (np.isclose(loss.item(), ref["loss"], rtol=3e-4) and np.isclose(scalars[1].item(), ref["ratio"], rtol=2e-4)
              and rel < 3e-2 and cos > 0.9999)
        out[rank] = (bool(ok), loss.item(), ref["loss"], scalars[1].item(), ref["ratio"], float(rel), float(cos))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode,gamma,cg", [(2, 6.0, True), (1, 5.0, False), (0, 1e6, False)])
def test_sharded_against_the_oracle(mode, gamma, cg):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    n_loc = 2560                                   # 2 * 2560 / 128 = 40 row tiles per rank
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_oracle_worker, args=(world, _free_port(), n_loc, 128, mode, gamma, cg, out), nprocs=world, join=True)
        res = dict(out)
        assert len(res) == world and all(v[0] for v in res.values()), res


@pytest.mark.parametrize("mode,gamma,cg", [(0, 1e6, False), (2, 6.0, True)])
def test_sharded_equals_single_gpu(mode, gamma, cg):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), 192, 128, mode, gamma, cg, out), nprocs=world, join=True)
        res = dict(out)
        assert all(v[0] for v in res.values()), res
