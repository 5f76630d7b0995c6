"""Hook-side glue of the contrastive pre-training step, without per-step host syncs.

Mirrors, for the fused loss only (paths under ``/root/reference``):

* label generators ``semi_seg/epochers/helper.py:48-65`` and their dispatch ``semi_seg/hooks/utils.py:10-65``
  -- same class names / keyword arguments, but without ``sklearn.LabelEncoder`` (``encode_labels`` restates
  ``fit(v).transform(v)``: the rank of each value among the sorted distinct values);
* the projector heads ``contrastyou/projectors/heads.py:9-25, :28-40, :76-115`` with their normalise /
  pool+normalise tails replaced by the CUDA kernels of this package;
* the per-batch hooks ``semi_seg/hooks/infonce.py:145-195`` (``_INFONCEEpochHook``), ``:198-241``
  (``_INFONCEDenseHook``), ``:244-268`` (``_SPINFONCEEpochHook``) and the per-epoch gamma update
  ``:133-141`` (``SelfPacedINFONCEHook.__call__``).

What changes against the reference: labels travel as cached device ``int32`` tensors instead of python lists
re-uploaded every call; ``loss`` / ``sp_weight`` / ``age_param`` meters accumulate on the device and are read
once per epoch (the reference calls ``loss.item()`` every batch, infonce.py:183); the N x N figures come from
the loss module's lazy diagnostics, so nothing N x N is formed unless a figure is actually drawn.  The
trainer / epocher framework, feature extractor and TensorBoard writer stay the reference's own host code
(DESIGN.md section 8): these classes take the extracted feature maps as arguments.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

import torch
from torch import Tensor, nn

from . import ops
from .dense import DenseProjectionTail, point_coordinates
from .losses import SelfPacedSupConLoss, SupConLoss1
from .projectors import Normalize
from .schedule import PScheduler

__all__ = ["encode_labels", "PartitionLabelGenerator", "PatientLabelGenerator", "ACDCCycleGenerator",
           "SIMCLRGenerator", "global_label_generator", "get_label", "DeviceMeter", "LabelCache", "ProjectionHead",
           "DenseProjectionHead", "INFONCEEpochHook", "INFONCEDenseHook", "SPINFONCEEpochHook",
           "SelfPacedGammaSchedule"]


# --------------------------------------------------------------------------------------------------
# label generators (helper.py:48-65)
# --------------------------------------------------------------------------------------------------
def encode_labels(values: Sequence) -> List[int]:
    """``LabelEncoder().fit(values).transform(values).tolist()``: index into the sorted distinct values."""
    order = {v: k for k, v in enumerate(sorted(set(values)))}
    return [order[v] for v in values]


class PartitionLabelGenerator:
    def __call__(self, partition_list: List[str], **kwargs):
        return encode_labels(partition_list)


class PatientLabelGenerator:
    def __call__(self, patient_list: List[str], **kwargs):
        return encode_labels(patient_list)


class ACDCCycleGenerator:
    def __call__(self, experiment_list: List[str], **kwargs):
        return [0 if e == "00" else 1 for e in experiment_list]


class SIMCLRGenerator:
    def __call__(self, partition_list: List[str], **kwargs):
        return list(range(len(partition_list)))


_GENERATORS = {
    "acdc": {"partition": PartitionLabelGenerator, "patient": PatientLabelGenerator, "cycle": ACDCCycleGenerator,
             "self": SIMCLRGenerator},
    "prostate": {"partition": PartitionLabelGenerator, "patient": PatientLabelGenerator, "self": SIMCLRGenerator},
    "mmwhs": {"partition": PartitionLabelGenerator, "patient": PatientLabelGenerator, "self": SIMCLRGenerator},
}


def global_label_generator(dataset_name: str, contrast_on: str):
    """hooks/utils.py:10-42."""
    try:
        return _GENERATORS[dataset_name][contrast_on]()
    except KeyError:
        raise NotImplementedError((dataset_name, contrast_on)) from None


def get_label(contrast_on, data_name, partition_group, label_group):
    """hooks/utils.py:45-65 (``label_group`` entries look like ``patient003_00``)."""
    if data_name == "acdc":
        return global_label_generator("acdc", contrast_on)(
            partition_list=partition_group, patient_list=[p.split("_")[0] for p in label_group],
            experiment_list=[p.split("_")[1] for p in label_group])
    if data_name in ("prostate", "prostate_md"):
        return global_label_generator("prostate", contrast_on)(
            partition_list=partition_group, patient_list=[p.split("_")[0] for p in label_group])
    if data_name in ("mmwhsct", "mmwhsmr"):
        return global_label_generator("mmwhs", contrast_on)(partition_list=partition_group, patient_list=label_group)
    raise NotImplementedError(data_name)


class LabelCache:
    """python label list -> device int32 tensor, uploaded once per distinct list (the sampler repeats batch
    compositions; SimCLR labels depend on the batch size only)."""

    def __init__(self, capacity: int = 64):
        self._cap = capacity
        self._store: "OrderedDict[tuple, Tensor]" = OrderedDict()

    def __call__(self, labels, device) -> Tensor:
        if isinstance(labels, Tensor):
            # same equality classes as torch.eq on the tensor itself (floats are not truncated, wide int64 values
            # do not wrap): ops.label_codes
            return ops.label_codes(labels, labels.shape[0], device)
        key = (str(device), tuple(int(v) for v in labels))
        hit = self._store.get(key)
        if hit is None:
            hit = ops.label_codes(list(key[1]), len(key[1]), device)
            self._store[key] = hit
            if len(self._store) > self._cap:
                self._store.popitem(last=False)
        else:
            self._store.move_to_end(key)
        return hit


class DeviceMeter:
    """Average-value meter whose additions stay on the device; ``summary()`` is the only host read."""

    def __init__(self):
        self._sum: Optional[Tensor] = None
        self._host_sum = 0.0
        self._n = 0

    def add(self, value) -> None:
        if isinstance(value, Tensor):
            v = value.detach().float().reshape(())
            self._sum = v.clone() if self._sum is None else self._sum.add_(v)
        else:
            self._host_sum += float(value)
        self._n += 1

    def reset(self) -> None:
        self._sum, self._host_sum, self._n = None, 0.0, 0

    def summary(self) -> float:
        if self._n == 0:
            return float("nan")
        dev = 0.0 if self._sum is None else float(self._sum.item())
        return (dev + self._host_sum) / self._n


# --------------------------------------------------------------------------------------------------
# projector heads (heads.py) with the fused tails
# --------------------------------------------------------------------------------------------------
class ProjectionHead(nn.Module):
    """heads.py:76-93 + :9-25: pool -> flatten -> Linear (-> LeakyReLU -> Linear) -> Normalize (CUDA kernel)."""

    def __init__(self, *, input_dim: int, hidden_dim=256, output_dim: int, head_type: str = "mlp",
                 normalize: bool = True, pool_name="adaptive_avg", spatial_size=(1, 1)):
        super().__init__()
        assert pool_name in ("adaptive_avg", "adaptive_max")
        assert head_type in ("mlp", "linear"), head_type
        pool = (nn.AdaptiveAvgPool2d if pool_name == "adaptive_avg" else nn.AdaptiveMaxPool2d)(tuple(spatial_size))
        flat_dim = input_dim * spatial_size[0] * spatial_size[1]
        if head_type == "mlp":
            body = [nn.Linear(flat_dim, hidden_dim), nn.LeakyReLU(0.01, inplace=True), nn.Linear(hidden_dim, output_dim)]
        else:
            body = [nn.Linear(flat_dim, output_dim)]
        self._header = nn.Sequential(pool, nn.Flatten(1), *body, Normalize() if normalize else nn.Identity())

    def forward(self, features):
        return self._header(features)


class DenseProjectionHead(nn.Module):
    """heads.py:96-120 + :28-40: 1x1 convs, then pool + normalise.  ``forward(features)`` keeps the reference's
    ``[B, C, ph, pw]`` output; ``rows(features, points=...)`` is the fused tail (pool + normalise + gather /
    reshape in one pass) that the dense hook uses."""

    def __init__(self, *, input_dim: int, hidden_dim=128, output_dim: int, head_type: str = "mlp",
                 normalize: bool = True, pool_name="adaptive_avg", spatial_size=(16, 16)):
        super().__init__()
        assert head_type in ("mlp", "linear"), head_type
        if head_type == "mlp":
            self._projector = nn.Sequential(nn.Conv2d(input_dim, hidden_dim, 1, 1, 0), nn.LeakyReLU(0.01, inplace=True),
                                            nn.Conv2d(hidden_dim, output_dim, 1, 1, 0))
        else:
            self._projector = nn.Sequential(nn.Conv2d(input_dim, output_dim, 1, 1, 0))
        self._spatial_size = tuple(int(v) for v in spatial_size)
        self._tail = DenseProjectionTail(self._spatial_size, pool_name, normalize)

    def rows(self, features, points: Optional[Tensor] = None, point_nums: Optional[int] = None,
             seed: Optional[int] = None) -> Tensor:
        """Unit anchor rows ``[B * P, C]`` (all pooled pixels, or the sampled ones).

        With average pooling the LAST 1x1 convolution is moved behind the pooling: both are linear, so
        ``pool(W h + b) == W pool(h) + b`` exactly (rounding aside), and the convolution then runs on ``ph * pw`` (or
        only the ``P`` sampled) positions per image instead of ``H * W`` -- 125x fewer for the dense hook's 10 x 10
        grid on a 112 x 112 map, 2500x with 5 sampled points -- and the full-resolution ``[B, C, H, W]`` projector
        output never exists.  Order: hidden layers at full resolution -> pool (+ gather) kernel -> ``F.linear`` on the
        rows -> normalise kernel.  Max pooling does not commute with the convolution and keeps the reference's order
        (``commute_pooling=False`` forces that order for average pooling too: A/B and parity tests)."""
        if points is None and point_nums is not None:
            points = point_coordinates(features.shape[0], *self._spatial_size, point_nums, seed)
        last = self._projector[-1]
        if self.commute_pooling and self._tail._pool == "avg":
            hidden = self._projector[:-1](features) if len(self._projector) > 1 else features
            pooled = ops.dense_rows(hidden, self._spatial_size, points, normalize=False)
            z = nn.functional.linear(pooled, last.weight.flatten(1), last.bias)
            return ops.l2norm_fwd(z, 1, 1e-12)[0]
        return self._tail(self._projector(features), points=points)

    #: average pooling: run the last 1x1 convolution on the pooled rows (see ``rows``)
    commute_pooling = True

    def forward(self, features):
        b = features.shape[0]
        ph, pw = self._spatial_size
        rows = self.rows(features)
        return rows.reshape(b, ph, pw, rows.shape[1]).permute(0, 3, 1, 2)


# --------------------------------------------------------------------------------------------------
# per-batch hooks
# --------------------------------------------------------------------------------------------------
class INFONCEEpochHook:
    """infonce.py:145-195 for the encoder (global) contrast.  ``__call__(features_tf, tf_features, ...)`` takes
    the two views' feature maps (already transformed as at :176-178) and returns ``loss * weight``."""

    def __init__(self, *, name: str = "infonce", weight: float = 1.0, projector: nn.Module, criterion,
                 label_generator=None, figure_fn=None) -> None:
        self._name = name
        self._weight = weight
        self._projector = projector
        self._criterion = criterion
        self._label_generator = label_generator
        self._labels = LabelCache()
        self._figure_fn = figure_fn              # callable(tensor, tag) -- e.g. the reference's figure2board
        self._n = 0
        self.meters: Dict[str, DeviceMeter] = {"loss": DeviceMeter()}

    def _target(self, partition_group, label_group, n: int, device):
        if self._label_generator is None:
            return None
        return self._labels(self._label_generator(partition_group=partition_group, label_group=label_group), device)

    def _project(self, features_tf: Tensor, tf_features: Tensor, seed):
        z = self._projector(torch.cat([features_tf, tf_features], dim=0))
        return torch.chunk(z, 2)

    def _figures(self):
        if self._n == 0 and self._figure_fn is not None:          # :188-192, first batch of the epoch only
            for tag in ("pos_mask", "sim_exp", "sim_logits"):
                self._figure_fn(self._criterion.figure(tag), tag)        # sub-sampled above 1024 anchors

    def __call__(self, features_tf: Tensor, tf_features: Tensor, *, partition_group=None, label_group=None,
                 seed: Optional[int] = None, **kwargs) -> Tensor:
        z_a, z_b = self._project(features_tf, tf_features, seed)
        target = self._target(partition_group, label_group, z_a.shape[0], z_a.device)
        loss = self._criterion(z_a, z_b, target=target)
        self.meters["loss"].add(loss)                            # device-side; the reference syncs here (:183)
        self._figures()
        self._n += 1
        return loss * self._weight

    def summary(self) -> Dict[str, float]:
        return {k: m.summary() for k, m in self.meters.items()}


class INFONCEDenseHook(INFONCEEpochHook):
    """infonce.py:198-241: dense (decoder) contrast on ``point_nums`` sampled pixels per image, SimCLR labels
    (:217).  Needs a ``DenseProjectionHead`` of this module: both views use the same coordinates, the ones
    ``FixRandomSeed(seed)`` reproduces in the reference (:209-212)."""

    def __init__(self, *, point_nums: int = 5, **kwargs) -> None:
        super().__init__(**kwargs)
        self._point_nums = point_nums

    def _project(self, features_tf, tf_features, seed):
        b = features_tf.shape[0]
        ph, pw = self._projector._spatial_size
        pts = point_coordinates(b, ph, pw, self._point_nums, seed)
        rows = self._projector.rows(torch.cat([features_tf, tf_features], dim=0), points=torch.cat([pts, pts]))
        return torch.chunk(rows, 2)

    def _target(self, partition_group, label_group, n, device):
        return None                               # target=list(range(n)) is the identity mask (:217, :140-143)


class SPINFONCEEpochHook(INFONCEEpochHook):
    """infonce.py:244-268: adds the ``sp_weight`` (downgrade ratio) and ``age_param`` meters."""

    def __init__(self, **kwargs) -> None:
        super().__init__(**kwargs)
        self.meters["sp_weight"] = DeviceMeter()
        self.meters["age_param"] = DeviceMeter()

    def __call__(self, *args, **kwargs) -> Tensor:
        loss = super().__call__(*args, **kwargs)
        self.meters["sp_weight"].add(self._criterion._scalars[1])     # device scalar; no .item() per batch
        self.meters["age_param"].add(self._criterion.age_param)
        if self._n == 1 and self._figure_fn is not None:              # after the first batch (:264)
            self._figure_fn(self._criterion.figure("sp_mask"), "sp_mask")
        return loss


class SelfPacedGammaSchedule:
    """``SelfPacedINFONCEHook`` (:108-141) minus the trainer plumbing: owns the criterion and the ``PScheduler``;
    ``new_epoch()`` does what its ``__call__`` does once per epoch (read gamma, step, ``set_gamma``)."""

    def __init__(self, *, mode="soft", p=0.5, begin_value=1e6, end_value=1e6, correct_grad: bool = False,
                 max_epoch: int, **criterion_kwargs) -> None:
        self.scheduler = PScheduler(max_epoch=int(max_epoch), begin_value=float(begin_value),
                                    end_value=float(end_value), p=float(p))
        self.criterion = SelfPacedSupConLoss(weight_update=mode, correct_grad=correct_grad, **criterion_kwargs)

    def new_epoch(self) -> float:
        gamma = self.scheduler.value
        self.scheduler.step()
        self.criterion.set_gamma(gamma)
        return gamma


def make_criterion(self_paced: bool = False, **kwargs):
    """``init_criterion`` of the two hook factories (:92-94, :127-131)."""
    return SelfPacedSupConLoss(**kwargs) if self_paced else SupConLoss1(**kwargs)
