"""CPU oracle for the LatteCLIP loss head -- TEST INFRASTRUCTURE ONLY.

This package is a plain torch-on-CPU restatement of the reference algorithm on the
hot path (open_clip ``ClipLoss`` / ``gather_features`` and the inline prototype /
pseudo-label / EMA / memory-bank code of ``train_one_epoch_v2``).  It is the checker:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it.  Nothing under ``latteclip_b200/``
imports it, and the product path fails loudly when the CUDA extension is missing.

Parity pin: the reference ships no tests and no golden vectors (SURVEY.md section 0,
fact 2).  The oracle is pinned instead against outputs of the reference itself, run
unmodified in the build container (``tests/golden/make_golden.py`` drives the real
``open_clip.loss.ClipLoss``, ``training.train.compute_text_weights`` and one full
step of the real ``training.train.train_one_epoch_v2`` through mock model/data
objects) and committed as fixtures under ``tests/golden/``.  When ``/root/reference``
is present the tests additionally compare the oracle against the live reference.

Every function cites the reference file:line it follows.
"""

from .clip_loss import (  # noqa: F401
    clip_loss_reference,
    clip_loss_rank_block,
    clip_loss_all_ranks,
    gather_features_emulated,
)
from .prototypes import (  # noqa: F401
    build_classifier,
    pseudo_label,
    text_margins,
    mix_and_ema,
    update_bank,
    prototype_step,
)
from . import zero_shot  # noqa: F401,E402
from . import siglip  # noqa: F401,E402
from . import distill  # noqa: F401,E402
from . import accum  # noqa: F401,E402
