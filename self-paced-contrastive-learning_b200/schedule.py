"""gamma ("age") schedule of the self-paced loss.

Restates ``semi_seg/hooks/infonce.py::PScheduler`` (:34-53) without its ``deepclustering2`` base class:
gamma(e) = begin + (end - begin) * (e / max_epoch) ** p, stepped once per epoch by the hook
(``SelfPacedINFONCEHook.__call__`` :133-141) and pushed into the loss with ``set_gamma``.
"""
from __future__ import annotations

__all__ = ["PScheduler"]


class PScheduler:
    def __init__(self, max_epoch, begin_value=0.0, end_value=1.0, p=0.5):
        self.max_epoch = max_epoch
        self.begin_value = float(begin_value)
        self.end_value = float(end_value)
        self.epoch = 0
        self.p = p

    def step(self):
        self.epoch += 1

    @property
    def value(self):
        return self.get_lr(self.epoch)

    def get_lr(self, cur_epoch):
        return self.begin_value + (self.end_value - self.begin_value) * (cur_epoch / self.max_epoch) ** self.p

    def state_dict(self):
        return dict(epoch=self.epoch)

    def load_state_dict(self, state):
        self.epoch = int(state["epoch"])
