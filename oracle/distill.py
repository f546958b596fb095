"""Oracle (TEST INFRASTRUCTURE ONLY): ``DistillClipLoss`` of open_clip, world size 1
(/root/reference/src/open_clip/loss.py:324-362), restated with materialised logits on the CPU.
Pinned by tests/golden/distill.npz, recorded from the reference class itself."""

from __future__ import annotations

import torch


def dist_loss(teacher_logits: torch.Tensor, student_logits: torch.Tensor) -> torch.Tensor:
    """loss.py:326-327."""
    return -(teacher_logits.softmax(dim=1) * student_logits.log_softmax(dim=1)).sum(dim=1).mean(dim=0)


def distill_clip_loss(image_features, text_features, logit_scale, dist_image_features, dist_text_features,
                      dist_logit_scale):
    """loss.py:329-361 with world_size == 1: get_logits (:102-118, scale on the A operand) for both
    models, then the mean of the two directions (:353-356).  The contrastive term is 0 (:346-351)."""
    logits_per_image = logit_scale * image_features @ text_features.T             # :115
    logits_per_text = logit_scale * text_features @ image_features.T              # :116
    dist_per_image = dist_logit_scale * dist_image_features @ dist_text_features.T
    dist_per_text = dist_logit_scale * dist_text_features @ dist_image_features.T
    return (dist_loss(dist_per_image, logits_per_image) + dist_loss(dist_per_text, logits_per_text)) / 2
