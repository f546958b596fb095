"""CPU oracle for SigLipLoss -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Plain torch-on-CPU restatement of ``src/open_clip/loss.py:453-560``; pinned by
``tests/golden/siglip.npz`` (the reference's own ``SigLipLoss`` at world size 1 and with its ring
exchange on real gloo groups of 2, 3 and 4 ranks, recorded by ``tests/golden/make_golden.py``).
"""

from typing import List, Sequence, Tuple

import torch
import torch.nn.functional as F


def siglip_block_loss(image_features: torch.Tensor, text_features: torch.Tensor,
                      logit_scale: torch.Tensor, logit_bias, negative_only: bool = False) -> torch.Tensor:
    """loss.py:509-519 (`_loss`): logits = s * I @ T.T + b (:504-507), labels = 2*eye - 1, or all -1
    for a block of another rank's texts (:498-502), -logsigmoid(labels * logits).sum() / rows."""
    logits = logit_scale * image_features @ text_features.T
    if logit_bias is not None:
        logits = logits + logit_bias
    n = image_features.shape[0]
    labels = -torch.ones((n, n), dtype=image_features.dtype)
    if not negative_only:
        labels = 2 * torch.eye(n, dtype=image_features.dtype) + labels
    return -F.logsigmoid(labels * logits).sum() / n


def siglip_all_ranks(image_shards: Sequence[torch.Tensor], text_shards: Sequence[torch.Tensor],
                     logit_scale: float, logit_bias: float, dtype=torch.float64
                     ) -> Tuple[List[torch.Tensor], List[torch.Tensor], List[torch.Tensor],
                                List[torch.Tensor], List[torch.Tensor]]:
    """Every rank's loss and gradients of loss.py:521-558 without the ring: rank r adds the
    positive block with its own texts (:522) and one negative-only block per other rank's texts
    (:535-541, :549-556); the exchange's backward returns each text shard's gradient to its owner
    (:419-428), so d_txt of rank q sums over all ranks' losses.  logit_scale / logit_bias gradients
    are per rank (of that rank's loss).  Returns (loss[W], dI[W], dT[W], ds[W], db[W])."""
    world = len(image_shards)
    imgs = [x.detach().to(dtype).clone().requires_grad_(True) for x in image_shards]
    txts = [x.detach().to(dtype).clone().requires_grad_(True) for x in text_shards]
    scales = [torch.tensor(float(logit_scale), dtype=dtype, requires_grad=True) for _ in range(world)]
    biases = [torch.tensor(float(logit_bias), dtype=dtype, requires_grad=True) for _ in range(world)]
    losses = []
    for r in range(world):
        loss = siglip_block_loss(imgs[r], txts[r], scales[r], biases[r])
        for q in range(world):
            if q != r:
                loss = loss + siglip_block_loss(imgs[r], txts[q], scales[r], biases[r], negative_only=True)
        losses.append(loss)
    torch.stack(losses).sum().backward()
    return ([x.detach() for x in losses], [x.grad for x in imgs], [x.grad for x in txts],
            [x.grad for x in scales], [x.grad for x in biases])
