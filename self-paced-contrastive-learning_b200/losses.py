"""Drop-in loss modules with the reference's surface.

Mirrors ``contrastyou/losses/contrast_loss3.py``: ``SupConLoss1`` (:34-110) and
``SelfPacedSupConLoss`` (:113-222) -- same constructor arguments, same
``forward(proj_feat1, proj_feat2, target=None, mask=None)`` precedence (mask > target > SimCLR
identity, :128-143), ``set_gamma`` / ``age_param`` (:216-222), ``downgrade_ratio`` (:191) and the
diagnostic attributes the hooks read (``sim_exp``, ``sim_logits``, ``pos_mask``, ``neg_mask``,
``sp_mask``; ``semi_seg/hooks/infonce.py:185-187, :263``).  The arithmetic runs in the fused CUDA
kernels behind ``spcl::supcon_fwd`` / ``spcl::supcon_bwd``; nothing N x N is stored unless one of
the diagnostic attributes is actually read.

Extra keyword arguments (accepted through the reference's ``**kwargs``):

``precision``   "auto" (default) | "bf16" | "fp32".  bf16 = tcgen05 tensor-core kernels, fp32 = SIMT
                kernels that match the reference to fp32 rounding.  "auto" takes fp32 for the
                tri-state ``mask=`` form, for N < 1024 (the reference's own batch sizes) and for
                temperatures below 0.025 (which the tensor-core kernels reject), bf16 above.  Note for
                ``weight_update="hard"``: with bf16 operands a pair whose l_ij lies within operand rounding
                (~2**-8 / T) of gamma can land on the other side of the threshold than in the fp32
                reference; ask for ``precision="fp32"`` when that must not happen.
``check_nan``   True (default) raises ``RuntimeError`` on a NaN loss right away like the reference
                (:203-204), which costs one host sync per call; False leaves the check to the caller.
``validate``    True (default) keeps the reference's ``assert is_normalized`` (:154), active only
                when Python runs without ``-O``, exactly like the reference.
``cuda_graph``  False (default).  True: forward + backward of a call run as ONE CUDA-graph replay (captured once
                per batch shape / gamma; labels form only) -- the reference's batch sizes are launch-bound.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor, nn

from . import _native as nat
from . import ops

__all__ = ["SupConLoss1", "SelfPacedSupConLoss", "SupConLoss2", "SupConLoss3", "SupConLoss4", "supcon_loss",
           "is_normalized", "grouped_forward"]

_AUTO_TC_MIN_N = 1024
_DIAG_MAX_N = 16384


def is_normalized(feature: Tensor, dim: int = 1) -> bool:
    """contrast_loss3.py:20-22."""
    norms = feature.norm(dim=dim)
    return bool(torch.allclose(norms, torch.ones_like(norms)))


#: the tensor-core kernels keep exp(S - 1/T) a normal fp32 only for 1/T <= 40 (supcon_tc.cu fill_params)
_TC_MAX_INV_TAU = 40.0


def _pick_tc(precision: str, N: int, has_tri: bool, mode: int = nat.MODE_NONE, temperature: float = 0.07) -> bool:
    """True = tcgen05 bf16 kernels, False = fp32 SIMT kernels.  ``"auto"`` never picks a path that would reject the
    call: tri-state masks, ``exclude_other_pos`` and temperatures below 0.025 stay on the fp32 kernels, which carry
    them at any N (the reference computes a loss for all of them)."""
    if mode == nat.MODE_EXCL:
        # exclude_other_pos (:97-100) has a per-pair denominator; it is carried by the fp32 kernels only
        if precision == "bf16":
            raise nat.SpclError("exclude_other_pos=True runs on the fp32 path only (precision='fp32' or 'auto')")
        return False
    if precision == "bf16":
        return True
    if precision == "fp32":
        return False
    if precision != "auto":
        raise ValueError(f"precision must be 'auto', 'bf16' or 'fp32', got {precision!r}")
    return (not has_tri) and N >= _AUTO_TC_MIN_N and 1.0 / float(temperature) <= _TC_MAX_INV_TAU


class _Diagnostics:
    """Lazily materialised N x N views of the last call (what the reference stashes at :175-178, :188)."""

    def __init__(self, z1, z2, labels, tri, row_stats, temperature, gamma, mode):
        self.z1, self.z2, self.labels, self.tri = z1.detach(), z2.detach(), labels, tri
        self.row_stats, self.t, self.gamma, self.mode = row_stats, temperature, gamma, mode
        self._cache = {}

    def subsampled(self, name, max_side):
        """``name`` restricted to ``max_side`` evenly spaced anchors (same rows and columns): what a figure of an
        N x N matrix can show anyway, for any N.  ``sim_logits`` is shifted by the sub-sample's own maximum -- the
        diagonal 1/T for unit rows, i.e. the reference's global maximum (:28-29)."""
        n = self.z1.shape[0]
        N = 2 * n
        if N <= max_side:
            return self.get(name)
        idx = torch.linspace(0, N - 1, max_side, device=self.z1.device).round().long().unique()
        z = torch.cat([self.z1, self.z2]).float()[idx]
        raw = (z @ z.t()) / self.t
        keep = 1.0 - torch.eye(idx.numel(), device=z.device)
        if name == "sim_logits":
            return raw - raw.max()
        if name == "sim_exp":
            return torch.exp(raw - raw.max())
        h = idx % n
        if self.tri is not None:
            t = self.tri[h][:, h]
            pos, neg = (t == 1).float() * keep, (t == 0).float() * keep
        else:
            lab = self.labels[h]
            same = lab[:, None] == lab[None, :]
            pos, neg = same.float() * keep, (~same).float() * keep
        if name == "pos_mask":
            return pos
        if name == "neg_mask":
            return neg
        if name == "sp_mask":
            l = self.row_stats[0, idx, None] - raw
            if self.mode in (nat.MODE_NONE, nat.MODE_EXCL):
                w = torch.ones_like(l)
            elif self.mode == nat.MODE_HARD:
                w = (l <= self.gamma).float()
            else:
                w = torch.clamp_min(1 - l / self.gamma, 0)
            return torch.maximum(w, 1 - pos)
        raise AttributeError(name)

    def _base(self):
        if "logits" not in self._cache:
            n = self.z1.shape[0]
            if 2 * n > _DIAG_MAX_N:
                raise RuntimeError(f"diagnostic N x N matrices are only materialised for N <= {_DIAG_MAX_N}")
            z = torch.cat([self.z1, self.z2]).float()
            logits = (z @ z.t()) / self.t
            logits = logits - logits.max()
            keep = 1.0 - torch.eye(2 * n, device=z.device)
            if self.tri is not None:
                t = self.tri.repeat(2, 2)
                pos, neg = (t == 1).float() * keep, (t == 0).float() * keep
            else:
                lab = torch.cat([self.labels, self.labels])
                same = lab[:, None] == lab[None, :]
                pos, neg = same.float() * keep, (~same).float() * keep
            self._cache.update(logits=logits, pos=pos, neg=neg)
        return self._cache

    def get(self, name):
        c = self._base()
        if name == "sim_logits":
            return c["logits"]
        if name == "sim_exp":
            return torch.exp(c["logits"])
        if name == "pos_mask":
            return c["pos"]
        if name == "neg_mask":
            return c["neg"]
        if name == "sp_mask":
            N = c["pos"].shape[0]
            z = torch.cat([self.z1, self.z2]).float()
            llh = (z @ z.t()) / self.t - self.row_stats[0, :N, None]
            l = -llh
            if self.mode in (nat.MODE_NONE, nat.MODE_EXCL):
                w = torch.ones_like(l)
            elif self.mode == nat.MODE_HARD:
                w = (l <= self.gamma).float()
            else:
                w = torch.clamp_min(1 - l / self.gamma, 0)
            return torch.maximum(w, 1 - c["pos"])
        raise AttributeError(name)


def supcon_loss(proj_feat1: Tensor, proj_feat2: Tensor, *, target=None, mask: Optional[Tensor] = None,
                temperature: float = 0.07, gamma: float = 1e6, mode: int = nat.MODE_NONE,
                correct_grad: bool = False, precision: str = "auto", graph_cache: Optional[dict] = None):
    """Functional form.  -> (loss 0-d, scalars[4] = loss/ratio/scale/scale_over_N, aux dict).

    ``graph_cache``: a dict owned by the caller; when given (labels form only), forward + backward of this call
    run as one CUDA-graph replay (``ops.GraphRunner``), captured once per (n, d, hyper-parameters)."""
    if proj_feat1.shape != proj_feat2.shape:
        raise AssertionError((proj_feat1.shape, proj_feat2.shape))
    if not (proj_feat1.is_cuda and proj_feat2.is_cuda):
        raise RuntimeError("spcl_b200 runs on CUDA tensors only: there is no CPU path (got "
                           f"{proj_feat1.device} / {proj_feat2.device})")
    n = proj_feat1.shape[0]
    dev = proj_feat2.device
    z1 = proj_feat1.float()
    z2 = proj_feat2.float()
    labels = tri = None
    if mask is not None:                      # :128-131
        tri = ops.tri_codes(mask, n, dev)
    elif target is not None:                  # :133-139
        labels = ops.label_codes(target, n, dev)
    else:                                     # :140-143  SimCLR
        labels = torch.arange(n, dtype=torch.int32, device=dev)
    use_tc = _pick_tc(precision, 2 * n, tri is not None, int(mode), temperature)
    if graph_cache is not None and tri is None and int(mode) != nat.MODE_EXCL and not torch.compiler.is_compiling():
        key = (n, z1.shape[1], str(dev), float(temperature), float(gamma), int(mode), bool(correct_grad), use_tc)
        runner = graph_cache.get(key)
        if runner is None:
            if len(graph_cache) >= 8:             # gamma changes once per epoch: keep the cache small
                graph_cache.pop(next(iter(graph_cache)))
                _note_graph_eviction()
            runner = graph_cache[key] = ops.GraphRunner(n, z1.shape[1], dev, temperature, gamma, mode, correct_grad,
                                                        use_tc)
        scalars, row_stats = ops.supcon_fwd_graphed(z1, z2, labels, runner)
    else:
        scalars, row_stats = ops.supcon_fwd_eager(z1, z2, labels, tri, float(temperature), float(gamma), int(mode),
                                                  bool(correct_grad), use_tc)
    return scalars[0], scalars, dict(labels=labels, tri=tri, row_stats=row_stats, use_tc=use_tc)


_GRAPH_EVICTIONS = [0]


def _note_graph_eviction():
    """gamma (and every other hyper-parameter) is baked into a captured graph: a caller that changes gamma on every
    step instead of once per epoch (infonce.py:134-136) recaptures a graph per call, which is slower than eager."""
    _GRAPH_EVICTIONS[0] += 1
    if _GRAPH_EVICTIONS[0] == 32:
        import warnings
        warnings.warn("spcl_b200: cuda_graph=True keeps recapturing graphs (32 evictions): the captured graph is keyed by "
                      "(batch shape, temperature, gamma, mode); change gamma once per epoch or use cuda_graph=False",
                      RuntimeWarning, stacklevel=3)


class _FusedSupConBase(nn.Module):
    _mode = nat.MODE_NONE

    def _init_common(self, temperature, kwargs):
        self._t = temperature
        self._precision = kwargs.pop("precision", "auto")
        self._check_nan = bool(kwargs.pop("check_nan", True))
        self._validate = bool(kwargs.pop("validate", True))
        self._graphs = {} if bool(kwargs.pop("cuda_graph", False)) else None
        self._diag = None
        self._scalars = None
        self._ratio_cache = None

    def _gamma_mode_cg(self):
        raise NotImplementedError

    def forward(self, proj_feat1, proj_feat2, target=None, mask: Tensor = None, **kwargs):
        if mask is not None:
            assert mask.shape == torch.Size([proj_feat1.size(0)] * 2), mask.shape
        if self._validate:
            assert is_normalized(proj_feat1) and is_normalized(proj_feat2), "features need to be normalized first"
        assert proj_feat1.shape == proj_feat2.shape, (proj_feat1.shape, proj_feat2.shape)
        gamma, mode, cg = self._gamma_mode_cg()
        loss, scalars, aux = supcon_loss(proj_feat1, proj_feat2, target=target, mask=mask, temperature=self._t,
                                         gamma=gamma, mode=mode, correct_grad=cg, precision=self._precision,
                                         graph_cache=self._graphs)
        # (plain attributes: nn.Module.__setattr__ checks every assignment against parameters / buffers / modules,
        # ~5 us each on the host -- more than a kernel launch at the reference's batch sizes)
        state = self.__dict__
        state["_scalars"] = scalars.detach()
        state["_ratio_cache"] = None
        state["_diag"] = _Diagnostics(proj_feat1, proj_feat2, aux["labels"], aux["tri"], aux["row_stats"], self._t,
                                      gamma, mode)
        if self._check_nan and torch.isnan(loss):      # :203-204 (one host sync, as in the reference)
            raise RuntimeError(loss)
        return loss

    def forward_raw(self, feat1: Tensor, feat2: Tensor, target=None, eps: float = 1e-12):
        """Loss from UN-normalised projector outputs ``[B, C]`` or ``[B, C, H, W]``: equals
        ``forward(rows(F.normalize(feat1, dim=1)), rows(F.normalize(feat2, dim=1)), target)`` with ``rows`` the
        reference's ``[b, c, h, w] -> [b*hw, c]`` reshape (comparable.py:398-404); ``target`` holds one label per
        anchor (``B`` or ``B*H*W`` entries).  On the tensor-core path the normalise, the reshape copy, the concat
        and the bf16 pack are one kernel and its backward consumes the loss-gradient rows directly; below the
        tensor-core threshold (and for ``exclude_other_pos``) it runs the unfused sequence."""
        assert feat1.shape == feat2.shape and feat1.dim() >= 2, (feat1.shape, feat2.shape)
        if not (feat1.is_cuda and feat2.is_cuda):
            raise RuntimeError("spcl_b200 runs on CUDA tensors only: there is no CPU path")
        outer, d = feat1.shape[0], feat1.shape[1]
        n = feat1.numel() // d
        gamma, mode, cg = self._gamma_mode_cg()
        if not _pick_tc(self._precision, 2 * n, False, int(mode), self._t):
            def rows(x):
                y = ops.l2norm_fwd(x.float(), 1, eps)[0]
                return y.reshape(outer, d, -1).permute(0, 2, 1).reshape(n, d)
            validate, self._validate = self._validate, False       # normalised by construction
            try:
                return self.forward(rows(feat1), rows(feat2), target=target)
            finally:
                self._validate = validate
        dev = feat1.device
        labels = ops.label_codes(target, n, dev) if target is not None else None
        scalars, row_stats = ops.supcon_fwd_raw(feat1.float(), feat2.float(), labels, float(self._t), float(gamma),
                                                int(mode), bool(cg), float(eps))
        self._scalars = scalars.detach()
        self._ratio_cache = None
        self._diag = None                       # the N x N diagnostics need the normalised rows: use forward()
        loss = scalars[0]
        if self._check_nan and torch.isnan(loss):
            raise RuntimeError(loss)
        return loss

    # diagnostics the hooks read after every call; materialised only on access
    def _diag_get(self, name):
        if self._diag is None:
            raise AttributeError(f"{name} is only available after a forward call")
        return self._diag.get(name)

    def figure(self, name: str, max_side: int = 1024) -> Tensor:
        """One of ``sim_exp / sim_logits / pos_mask / neg_mask / sp_mask`` of the last call for plotting: the full
        matrix up to ``max_side`` anchors, an evenly sub-sampled one above (works at any N; the attribute forms
        below materialise the full N x N matrix like the reference and refuse beyond N = 16384)."""
        if self._diag is None:
            raise AttributeError(f"{name} is only available after a forward call")
        return self._diag.subsampled(name, int(max_side))

    sim_exp = property(lambda self: self._diag_get("sim_exp"))
    sim_logits = property(lambda self: self._diag_get("sim_logits"))
    pos_mask = property(lambda self: self._diag_get("pos_mask"))
    neg_mask = property(lambda self: self._diag_get("neg_mask"))


class SupConLoss1(_FusedSupConBase):
    """contrast_loss3.py:34-110 (W == 1)."""

    def __init__(self, temperature=0.07, exclude_other_pos=False, **kwargs):
        super().__init__()
        self._init_common(temperature, kwargs)
        self._exclude_pos = bool(exclude_other_pos)      # :97-100 (fp32 kernels, any N)

    def _gamma_mode_cg(self):
        return 1e6, (nat.MODE_EXCL if self._exclude_pos else nat.MODE_NONE), False


class SelfPacedSupConLoss(_FusedSupConBase):
    """contrast_loss3.py:113-222."""

    def __repr__(self):
        return f"{self.__class__.__name__} with T: {self._t}, method: {self._weight_update} gamma: {self.__gamma}"

    def __init__(self, temperature=0.07, weight_update="hard", correct_grad=False, **kwargs):
        super().__init__()
        self._init_common(temperature, kwargs)
        self._weight_update = weight_update
        self.__gamma = 1e6
        self._correct_grad = correct_grad

    def _gamma_mode_cg(self):
        # every value other than "hard" selects the soft rule in the reference (:210-213)
        mode = nat.MODE_HARD if self._weight_update == "hard" else nat.MODE_SOFT
        return self.__gamma, mode, bool(self._correct_grad)

    def set_gamma(self, gamma):
        self.__gamma = float(gamma)

    @property
    def age_param(self):
        return self.__gamma

    @property
    def downgrade_ratio(self):
        """mean of W over the positive pairs of the last call (:189-191); read lazily (one sync)."""
        if self._scalars is None:
            raise AttributeError("downgrade_ratio is only available after a forward call")
        if self._ratio_cache is None:
            self._ratio_cache = float(self._scalars[1].item())
        return self._ratio_cache

    sp_mask = property(lambda self: self._diag_get("sp_mask"))


# ----------------------------------------------------------------------------------------------------------------
# soft positive weights: the older loss generation, contrastyou/losses/contrast_loss.py
# ----------------------------------------------------------------------------------------------------------------
class _WeightedBase(nn.Module):
    """Shared surface of ``SupConLoss2/3/4`` (contrast_loss.py:33-270): ``__init__(temperature=0.07, out_mode=True)``,
    ``RuntimeError`` on a NaN loss (:98-99), ``sim_exp`` / ``sim_logits`` / ``pos_weight`` kept lazily for the
    ``register_writer`` figures.  The arithmetic runs in ``spcl_supcon_fwd_w_f32`` / ``_bwd_w_f32`` (fp32 kernels:
    these losses carry an N x N weight matrix by nature and run at the reference's batch sizes)."""

    def __init__(self, temperature=0.07, out_mode=True, **kwargs):
        super().__init__()
        self._t = temperature
        self._out_mode = out_mode
        self._check_nan = bool(kwargs.pop("check_nan", True))
        self._validate = bool(kwargs.pop("validate", True))
        self._last = None

    def _run(self, proj_feat1, proj_feat2, pw, enable):
        if self._validate:
            assert is_normalized(proj_feat1) and is_normalized(proj_feat2), "features need to be normalized first"
        assert proj_feat1.shape == proj_feat2.shape, (proj_feat1.shape, proj_feat2.shape)
        if not (proj_feat1.is_cuda and proj_feat2.is_cuda):
            raise RuntimeError("spcl_b200 runs on CUDA tensors only: there is no CPU path")
        loss = ops.supcon_weighted(proj_feat1, proj_feat2, pw, enable, self._t, not self._out_mode)
        self._last = (proj_feat1.detach(), proj_feat2.detach())
        if self._check_nan and torch.isnan(loss):
            raise RuntimeError(loss)
        return loss

    def _logits(self):
        if self._last is None:
            raise AttributeError("only available after a forward call")
        z = torch.cat(self._last).float()
        logits = (z @ z.t()) / self._t
        return logits - logits.max()

    sim_logits = property(lambda self: self._logits())
    sim_exp = property(lambda self: torch.exp(self._logits()))


class SupConLoss2(_WeightedBase):
    """contrast_loss.py:33-100: label / tri-state ``mask`` form with the "in" (``out_mode=False``, :90-92) or "out"
    (:94-97) placement of the logarithm."""

    def forward(self, proj_feat1, proj_feat2, target=None, mask: Tensor = None):
        if (target is not None) and (mask is not None):
            raise RuntimeError("`target` and `mask` should not be provided in the same time")      # :52-53
        n, dev = proj_feat1.shape[0], proj_feat2.device
        if mask is not None:
            assert mask.shape == torch.Size([n, n])
            m = mask.to(dev)
            pos, neg = (m == 1), (m == 0)
        elif target is not None:
            codes = ops.label_codes(target, n, dev)
            pos = codes[:, None] == codes[None, :]
            neg = ~pos
        else:
            pos = torch.eye(n, dtype=torch.bool, device=dev)
            neg = ~pos
        self.pos_mask, self.neg_mask = pos.repeat(2, 2), neg.repeat(2, 2)
        enable = (pos | neg).repeat(2, 2).to(torch.uint8)             # pairs of the denominator: pos + negs (:85-92)
        return self._run(proj_feat1, proj_feat2, pos.float(), enable)


class SupConLoss3(_WeightedBase):
    """contrast_loss.py:130-182 ("soften supervised contrastive loss"): ``pos_weight`` [n, n], tiled 2 x 2 (:152)."""

    def forward(self, proj_feat1, proj_feat2, pos_weight: Tensor = None, **kwargs):
        assert pos_weight is not None
        n = len(proj_feat1)
        assert pos_weight.shape == torch.Size([n, n])
        pw = pos_weight.detach().to(proj_feat2.device, torch.float32)
        self.pos_weight = pw.repeat(2, 2)
        return self._run(proj_feat1, proj_feat2, pw, None)


class SupConLoss4(_WeightedBase):
    """contrast_loss.py:206-262: separate weight blocks within view 1, within view 2 and across the views; pairs of a
    block that was not given are left out of the denominator (``enable_mask``, :246).  The reference fills the
    view-1 block only when ``one2two_weight`` is given (:217-219); that is kept."""

    def forward(self, *, proj_feat1, proj_feat2, one2one_weight: Tensor = None, two2two_weight: Tensor,  # noqa
                one2two_weight=None, **kwargs):  # noqa
        assert one2one_weight is not None or one2two_weight is not None or two2two_weight is not None
        n, dev = len(proj_feat1), proj_feat2.device
        pw = torch.zeros(2 * n, 2 * n, device=dev, dtype=torch.float32)
        en = torch.zeros(2 * n, 2 * n, device=dev, dtype=torch.uint8)
        if one2two_weight is not None:                              # (sic) :217
            pw[:n, :n] = one2one_weight
            en[:n, :n] = 1
        if two2two_weight is not None:
            pw[n:, n:] = two2two_weight
            en[n:, n:] = 1
        if one2two_weight is not None:
            pw[:n, n:] = one2two_weight
            pw[n:, :n] = one2two_weight
            en[:n, n:] = 1
            en[n:, :n] = 1
        self.pos_weight, self.enable_mask = pw, en.float()
        return self._run(proj_feat1, proj_feat2, pw, en)


_GROUP_GRAPHS: dict = {}


def grouped_forward(criteria, feats, targets=None, cuda_graph: bool = False):
    """The K contrastive losses of one training step -- one criterion, projector output pair and label list per
    meta-label (partition / patient / cycle; poster Eq. 4, ``semi_seg`` ``creator.py:102-124``) -- evaluated with ONE
    kernel launch per stage for the whole group instead of K separate forward calls.  Equivalent to
    ``[c(z1, z2, target=t) for c, (z1, z2), t in zip(criteria, feats, targets)]`` (same losses, gradients,
    ``downgrade_ratio``); problems the grouped fp32 kernels do not carry (tensor-core sizes, ``exclude_other_pos``)
    make the whole group run as those separate calls.  The N x N diagnostics are not kept for grouped calls.
    ``cuda_graph=True`` replays forward + backward of the whole group as one captured graph (per shapes / gammas)."""
    K = len(criteria)
    targets = [None] * K if targets is None else list(targets)
    if len(feats) != K or len(targets) != K:
        raise ValueError("criteria, feats and targets must have the same length")
    metas, labels, groupable = [], [], K <= nat.MAX_GROUP
    for crit, (z1, z2), t in zip(criteria, feats, targets):
        assert z1.shape == z2.shape, (z1.shape, z2.shape)
        if crit._validate:
            assert is_normalized(z1) and is_normalized(z2), "features need to be normalized first"
        gamma, mode, cg = crit._gamma_mode_cg()
        n = z1.shape[0]
        if int(mode) == nat.MODE_EXCL or _pick_tc(crit._precision, 2 * n, False, int(mode), crit._t):
            groupable = False
        metas.append((crit._t, gamma, mode, cg))
    if not groupable:
        return [c(z1, z2, target=t) for c, (z1, z2), t in zip(criteria, feats, targets)]
    for (z1, _), t in zip(feats, targets):
        n, dev = z1.shape[0], z1.device
        lab = ops.label_codes(t, n, dev) if t is not None else torch.arange(n, dtype=torch.int32, device=dev)
        labels.append(lab if cuda_graph else torch.cat([lab, lab]))
    if cuda_graph:
        shapes = [tuple(z1.shape) for z1, _ in feats]
        key = (str(feats[0][0].device), tuple(shapes), tuple((float(t), float(g), int(m), bool(c)) for t, g, m, c in metas))
        runner = _GROUP_GRAPHS.get(key)
        if runner is None:
            if len(_GROUP_GRAPHS) >= 8:               # gammas change once per epoch: keep the cache small
                _GROUP_GRAPHS.pop(next(iter(_GROUP_GRAPHS)))
            runner = _GROUP_GRAPHS[key] = ops.GroupGraphRunner(shapes, metas, feats[0][0].device)
        scalars = list(ops.supcon_group_f32_graphed(list(feats), labels, runner).unbind(0))
    else:
        scalars = ops.supcon_group_f32(list(feats), labels, metas)
    out = []
    for crit, sc in zip(criteria, scalars):
        state = crit.__dict__                    # (plain attributes, see forward)
        state["_scalars"], state["_ratio_cache"], state["_diag"] = sc.detach(), None, None
        loss = sc[0]
        if crit._check_nan and torch.isnan(loss):
            raise RuntimeError(loss)
        out.append(loss)
    return out
