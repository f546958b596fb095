"""latteclip_b200 -- B200-native (sm_100a) loss head for LatteCLIP.

Only the data-parallel hot path is here: open_clip's ClipLoss / gather_features /
create_loss (``latteclip_b200.loss``) and LatteCLIP's prototype / pseudo-label / mixture /
EMA / memory-bank path (``latteclip_b200.prototypes``) plus the zero-shot evaluation /
feature-record row next to it (``latteclip_b200.zero_shot``), all thin Python over a C-ABI CUDA
library (``latteclip_b200/csrc``, header ``include/latte_b200.h``).  Encoders, data
pipeline and training driver remain the reference's PyTorch code.
"""

from .loss import ClipLoss, create_loss, gather_features  # noqa: F401
from .siglip import SigLipLoss  # noqa: F401
from .distill import DistillClipLoss  # noqa: F401
from . import train_step  # noqa: F401
from . import prototypes  # noqa: F401
from . import zero_shot  # noqa: F401
from ._lib import build, version, clear_workspace_cache  # noqa: F401

__all__ = ["ClipLoss", "SigLipLoss", "DistillClipLoss", "train_step", "create_loss", "gather_features", "prototypes", "zero_shot", "build", "version",
           "clear_workspace_cache"]
