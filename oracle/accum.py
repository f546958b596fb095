"""Oracle (TEST INFRASTRUCTURE ONLY): the ``--accum-freq > 1`` feature-cache path of upstream
open_clip as it stands in the reference behind ``raise NotImplemented()``
(/root/reference/src/training/train.py:972-1024).

For micro-batch j of one accumulation cycle the loss is ``ClipLoss`` over the CONCATENATION of all
micro-batches' features, where block j carries gradient (re-computed with grad) and the other
blocks are the cached no-grad features (train.py:1009-1016); ``backward`` runs once per micro-batch
(train.py:1023), so logit_scale receives the full-batch gradient ``accum_freq`` times.
"""

from __future__ import annotations

from typing import List, Sequence

import torch

from .clip_loss import clip_loss_reference


def clip_loss_accumulated(img_blocks: Sequence[torch.Tensor], txt_blocks: Sequence[torch.Tensor],
                          logit_scale: float, dtype: torch.dtype = torch.float64):
    """-> per micro-batch j: (loss_j, dI_j [m, D], dT_j [m, D], ds_j) exactly as train.py:1002-1023
    computes them (world size 1)."""
    out = []
    k = len(img_blocks)
    cached_i = [x.detach().to(dtype) for x in img_blocks]
    cached_t = [x.detach().to(dtype) for x in txt_blocks]
    for j in range(k):
        live_i = cached_i[j].clone().requires_grad_(True)
        live_t = cached_t[j].clone().requires_grad_(True)
        s = torch.tensor(float(logit_scale), dtype=dtype, requires_grad=True)
        all_i = torch.cat(cached_i[:j] + [live_i] + cached_i[j + 1:])        # train.py:1013-1015
        all_t = torch.cat(cached_t[:j] + [live_t] + cached_t[j + 1:])
        loss = clip_loss_reference(all_i, all_t, s)                           # train.py:1017
        loss.backward()                                                       # train.py:1023
        out.append((loss.detach(), live_i.grad, live_t.grad, s.grad))
    return out
