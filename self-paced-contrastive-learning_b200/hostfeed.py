"""Host-buffer front end of the fused loss: pinned-host embeddings in, loss (and gradients) out.

The reference's callers hand the loss device tensors (``semi_seg/hooks/infonce.py:182``); a caller that keeps
its projector outputs in host memory (the C-ABI / plugin use of ``include/spcl.h``, ``bench.py``'s ``e2e`` leg)
pays a PCIe copy of ``2 n d`` floats per step.  ``HostFeed`` takes that copy off the critical path: a ring of
``depth`` device staging slots is filled on a side stream, so the host->device copy of batch ``k + 1`` runs
while batch ``k`` is inside the fused kernels.  Nothing is cached between steps -- every batch is copied, every
loss is computed and read back.

    feed = HostFeed(n, d, device)
    feed.push(z1_host, z2_host, labels_host)              # batch 0
    for k in range(steps):
        if k + 1 < steps:
            feed.push(*next_batch)                         # overlaps with the compute below
        z1, z2, labels, slot = feed.pop()                  # current stream waits for that slot's copy only
        loss = criterion(z1, z2, target=labels); loss.backward()
        feed.release(slot, loss)                           # slot reusable once this step is done; loss -> host ring
    losses = feed.losses()                                 # one synchronisation at the end
"""
from __future__ import annotations

import collections
from typing import List, Optional, Tuple

import torch
from torch import Tensor

__all__ = ["HostFeed"]


class _Slot:
    def __init__(self, n, d, device, label_dtype):
        self.z1 = torch.empty(n, d, dtype=torch.float32, device=device)
        self.z2 = torch.empty(n, d, dtype=torch.float32, device=device)
        self.labels = torch.empty(n, dtype=label_dtype, device=device)
        self.ready = torch.cuda.Event()
        self.free = torch.cuda.Event()
        self.used = False


class HostFeed:
    def __init__(self, n: int, d: int, device, depth: int = 2, label_dtype=torch.int32, loss_ring: int = 4096):
        if depth < 2:
            raise ValueError("depth must be >= 2 (one slot in the kernels, one being filled)")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostFeed stages into CUDA memory; there is no CPU path")
        self.n, self.d = n, d
        self._slots = [_Slot(n, d, self.device, label_dtype) for _ in range(depth)]
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._next_push = 0
        self._queue = collections.deque()
        self._loss_host = torch.empty(loss_ring, dtype=torch.float32).pin_memory()
        self._n_loss = 0
        self.h2d_bytes = 0

    def push(self, z1_host: Tensor, z2_host: Tensor, labels_host: Optional[Tensor] = None) -> None:
        """Enqueue the host->device copy of one batch (pinned host tensors give a truly asynchronous copy)."""
        if len(self._queue) == len(self._slots):
            raise RuntimeError("every staging slot is in flight: pop()/release() one before the next push()")
        if tuple(z1_host.shape) != (self.n, self.d) or tuple(z2_host.shape) != (self.n, self.d):
            raise AssertionError((tuple(z1_host.shape), tuple(z2_host.shape), (self.n, self.d)))
        idx = self._next_push
        slot = self._slots[idx]
        self._next_push = (idx + 1) % len(self._slots)
        cs = self._copy_stream
        if slot.used:
            cs.wait_event(slot.free)                       # the step that last read this slot has finished
        with torch.cuda.stream(cs):
            slot.z1.copy_(z1_host, non_blocking=True)
            slot.z2.copy_(z2_host, non_blocking=True)
            self.h2d_bytes += 2 * self.n * self.d * 4
            if labels_host is not None:
                slot.labels.copy_(labels_host, non_blocking=True)
                self.h2d_bytes += labels_host.numel() * labels_host.element_size()
            slot.ready.record(cs)
        slot.used = True
        self._queue.append((idx, labels_host is not None))

    def pop(self) -> Tuple[Tensor, Tensor, Optional[Tensor], int]:
        """-> (z1, z2, labels | None, slot id): fresh autograd leaves over the oldest staged batch."""
        idx, has_labels = self._queue.popleft()
        slot = self._slots[idx]
        torch.cuda.current_stream(self.device).wait_event(slot.ready)
        z1 = slot.z1.detach().requires_grad_(True)
        z2 = slot.z2.detach().requires_grad_(True)
        return z1, z2, (slot.labels if has_labels else None), idx

    def release(self, slot_id: int, loss: Optional[Tensor] = None) -> None:
        """Marks the slot reusable after the work queued so far; optionally starts the loss read-back."""
        if loss is not None:
            k = self._n_loss % self._loss_host.numel()
            self._loss_host[k:k + 1].copy_(loss.detach().reshape(1), non_blocking=True)
            self._n_loss += 1
        self._slots[slot_id].free.record(torch.cuda.current_stream(self.device))

    def losses(self) -> List[float]:
        """Synchronises and returns the losses read back since the last call."""
        torch.cuda.synchronize(self.device)
        m = min(self._n_loss, self._loss_host.numel())
        out = self._loss_host[:m].tolist()
        self._n_loss = 0
        return out
