"""spcl_b200 -- B200-native fused self-paced supervised-contrastive loss.

One hot path of jizongFox/Self-paced-Contrastive-Learning (``contrastyou/losses/contrast_loss3.py``
and the projector's L2-normalise tail) as hand-written sm_100a CUDA kernels behind a C ABI
(``include/spcl.h``) and torch custom ops.  See DESIGN.md / INTEGRATION.md.
"""
from . import _native
from ._native import LIB_PATH, SpclError, build
from . import dense, hooks
from .dense import DenseProjectionTail, point_coordinates, region_extractor
from .hostfeed import HostFeed
from .losses import (SelfPacedSupConLoss, SupConLoss1, SupConLoss2, SupConLoss3, SupConLoss4, grouped_forward,
                     is_normalized, supcon_loss)
from .projectors import Normalize, normalize
from .schedule import PScheduler

__all__ = ["SelfPacedSupConLoss", "SupConLoss1", "SupConLoss2", "SupConLoss3", "SupConLoss4", "supcon_loss", "grouped_forward", "is_normalized", "Normalize", "normalize",
           "PScheduler", "HostFeed", "DenseProjectionTail", "point_coordinates", "region_extractor", "dense", "hooks", "build", "LIB_PATH", "SpclError"]
__version__ = "0.1.0"
