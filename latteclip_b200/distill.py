"""Drop-in for open_clip's ``DistillClipLoss`` (/root/reference/src/open_clip/loss.py:324-362).

    dist_loss(T, S) = -(softmax(T, dim=1) * log_softmax(S, dim=1)).sum(dim=1).mean(dim=0)     (:326-327)
    distill_loss    = (dist_loss(T_img, S_img) + dist_loss(T_txt, S_txt)) / 2                   (:353-356)

with S = s I T^T the student's logits and T the teacher's (the reference's contrastive term is
commented out, :346-351: the module returns 0 for it).  Neither [N, N] matrix is materialised: with
W(X) = softmax_rows(X) + softmax_cols(X) the loss is
``(sum_i lse_j S_ij + sum_j lse_i S_ij - s <I, W(T) T_s>) / (2N)`` and its gradient with respect to S is
``(W(S) - W(T)) / (2N)``, so forward and backward are the ClipLoss kernels run on two operand pairs:
forward sweeps for the LSE vectors of both models, the gradient sweep (without its label term) for
W, and the stream-K gradient GEMM for the products with the student's features
(include/latte_b200.h: latte_distill_products / latte_distill_loss / latte_distill_bwd_combine).

Scope: world size 1 (the reference gathers through ``get_logits``; that exchange is not built for
this loss), features of width <= 768 (multiple of 8) computed on the tensor cores in the autocast
dtype, or fp16 for fp32 inputs without autocast.  The teacher's features receive no gradient (the
reference evaluates the teacher under ``torch.no_grad()``, train.py:879-882).
"""

from __future__ import annotations

import torch

from . import _lib
from .loss import ClipLoss


class _FusedDistillLoss(torch.autograd.Function):

    @staticmethod
    def forward(ctx, image_features, text_features, logit_scale, dist_image_features,
                dist_text_features, dist_logit_scale, compute_dtype):
        img, _ = _lib.prep_features(image_features.detach(), compute_dtype)
        txt, _ = _lib.prep_features(text_features.detach(), compute_dtype)
        t_img, _ = _lib.prep_features(dist_image_features.detach(), compute_dtype)
        t_txt, _ = _lib.prep_features(dist_text_features.detach(), compute_dtype)
        row_s, col_s, _ = _lib.clip_fwd(img, txt, img, txt, 0, logit_scale)                      # loss.py:341-342
        row_t, col_t, _ = _lib.clip_fwd(t_img, t_txt, t_img, t_txt, 0, dist_logit_scale)         # :344-345
        a_t, b_t = _lib.distill_products(t_img, t_txt, img, txt, dist_logit_scale, row_t, col_t)
        loss, dot_t = _lib.distill_loss(row_s, col_s, img, a_t, logit_scale)                     # :353-356
        ctx.save_for_backward(img, txt, logit_scale.detach(), row_s, col_s, a_t, b_t, dot_t)
        ctx.meta = (image_features.dtype, text_features.dtype, logit_scale.dtype, logit_scale.shape)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        img, txt, scale, row_s, col_s, a_t, b_t, dot_t = ctx.saved_tensors
        di_t, dt_t, s_dtype, s_shape = ctx.meta
        gdt = di_t if di_t == dt_t and di_t in (torch.float32, torch.bfloat16, torch.float16) else torch.float32
        a_s, b_s = _lib.distill_products(img, txt, img, txt, scale, row_s, col_s)
        d_img, d_txt, d_scale = _lib.distill_bwd_combine(a_s, a_t, b_s, b_t, img, scale, grad_out, dot_t, gdt)
        need = ctx.needs_input_grad
        return (d_img.to(di_t) if need[0] else None, d_txt.to(dt_t) if need[1] else None,
                d_scale.reshape(s_shape).to(s_dtype) if need[2] else None, None, None, None, None)


class DistillClipLoss(ClipLoss):

    def dist_loss(self, teacher_logits, student_logits):
        # loss.py:326-327 -- materialising utility, NOT used by forward()
        return -(teacher_logits.softmax(dim=1) * student_logits.log_softmax(dim=1)).sum(dim=1).mean(dim=0)

    def forward(
            self,
            image_features,
            text_features,
            logit_scale,
            dist_image_features,
            dist_text_features,
            dist_logit_scale,
            output_dict=False,
    ):
        if self.world_size > 1:
            raise NotImplementedError("latteclip_b200.DistillClipLoss: world_size > 1 is not built")
        text_features = text_features.squeeze()                  # loss.py:338-339
        dist_text_features = dist_text_features.squeeze()
        if image_features.shape != text_features.shape or dist_image_features.shape != dist_text_features.shape \
                or dist_image_features.shape[0] != image_features.shape[0]:
            raise RuntimeError("DistillClipLoss: student and teacher need one batch size, and each model one "
                               "feature shape for both towers")
        if dist_image_features.shape[1] != image_features.shape[1]:
            raise NotImplementedError("latteclip_b200.DistillClipLoss: teacher and student widths differ")
        dev = image_features.device
        if not torch.is_tensor(logit_scale):
            logit_scale = torch.tensor(float(logit_scale), device=dev)
        if not torch.is_tensor(dist_logit_scale):
            dist_logit_scale = torch.tensor(float(dist_logit_scale), device=dev)
        if torch.is_autocast_enabled():
            cdt = torch.get_autocast_dtype("cuda")
        else:
            cdt = torch.promote_types(image_features.dtype, text_features.dtype)
        if cdt not in (torch.bfloat16, torch.float16):
            cdt = torch.float16      # fp32 inputs: the tensor-core operands are fp16 (11-bit mantissa)
        if not _lib.rank_sweep_supported(cdt, image_features.shape[1]):
            raise NotImplementedError("latteclip_b200.DistillClipLoss: feature width must be <= 768 and a "
                                      "multiple of 8")
        with torch.autocast(device_type="cuda", enabled=False):
            distill_loss = _FusedDistillLoss.apply(image_features, text_features, logit_scale,
                                                   dist_image_features, dist_text_features,
                                                   dist_logit_scale.detach(), cdt)
        if output_dict:
            return {"contrastive_loss": 0, "distill_loss": distill_loss}     # loss.py:358-359
        return 0, distill_loss                                               # :361
