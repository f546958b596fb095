"""SigLipLoss of open_clip (``src/open_clip/loss.py:453-560``) on the fused tile engine.

Same constructor and call signature as the reference.  Per rank the loss is
``sum_{i in own images, j in ALL texts} -logsigmoid(label_ij * (s <I_i, T_j> + b)) / n`` with
``label_ij = +1`` iff j is image i's own text (loss.py:498-519).  The reference reaches the other
ranks' texts by passing text shards round the ring and adding one negative-only block per hop
(:521-558); here the text shards are all-gathered once and one logit sweep covers every block.
The backward of the ring exchange returns each text shard's gradient to its owner (:419-428):
that is a reduce-scatter of the text-side product, done by NCCL or fused into the gradient GEMM
over peer-mapped accumulators exactly as for ClipLoss (``loss._peer_accumulator``).

16-bit features only (the reference's amp / amp_bf16 precision); no CPU or PyTorch fallback.
"""

from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .loss import _all_gather_cat, _peer_accumulator, _reduce_scatter_sum


class _FusedSigLip(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image_features, text_features, logit_scale, logit_bias, rank, world_size, group):
        img = image_features.detach().contiguous()
        txt = text_features.detach().contiguous()
        n = img.shape[0]
        if world_size > 1:
            all_txt = _all_gather_cat(txt, group)           # rank-major rows, as the ring visits them
            label_offset = rank * n
        else:
            all_txt, label_offset = txt, 0
        loss = _lib.siglip_fwd(img, all_txt, label_offset, logit_scale, logit_bias).reshape(())
        bias = logit_bias.detach() if logit_bias is not None else None
        ctx.save_for_backward(img, all_txt, logit_scale.detach(), bias)
        ctx.cfg = (rank, world_size, group, label_offset)
        ctx.meta = (logit_scale.dtype, logit_scale.shape,
                    None if logit_bias is None else (logit_bias.dtype, logit_bias.shape))
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        img, all_txt, scale, bias = ctx.saved_tensors
        rank, world_size, group, label_offset = ctx.cfg
        n = img.shape[0]
        if world_size == 1:
            d_img, d_txt, d_scale, d_bias = _lib.siglip_bwd(img, all_txt, 0, scale, bias, grad_out)
        else:
            peer = _peer_accumulator(n, img.shape[1], img.device, group)
            if peer is not None:
                acc, hdl, ptrs = peer
                acc.zero_()
                hdl.barrier(channel=0)        # all accumulators zeroed, last step's reads done
                d_img, _, d_scale, d_bias = _lib.siglip_bwd(img, all_txt, label_offset, scale, bias,
                                                            grad_out, peer_ptrs=ptrs)
                hdl.barrier(channel=1)        # every rank's adds have landed
                d_txt = acc.to(img.dtype)
            else:
                d_img, d_part, d_scale, d_bias = _lib.siglip_bwd(img, all_txt, label_offset, scale,
                                                                 bias, grad_out, partial=True)
                d_txt = _reduce_scatter_sum(d_part, n, rank, group).to(img.dtype)
        s_dtype, s_shape, b_meta = ctx.meta
        d_scale = d_scale.reshape(s_shape).to(s_dtype)
        d_bias = d_bias.reshape(b_meta[1]).to(b_meta[0]) if b_meta is not None else None
        need = ctx.needs_input_grad
        return (d_img if need[0] else None, d_txt if need[1] else None, d_scale if need[2] else None,
                d_bias if (b_meta is not None and need[3]) else None, None, None, None)


class SigLipLoss(nn.Module):
    """Drop-in for ``open_clip.loss.SigLipLoss`` (loss.py:453-560)."""

    def __init__(self, cache_labels=False, rank=0, world_size=1, bidir=True, use_horovod=False):
        super().__init__()
        self.cache_labels = cache_labels
        self.rank = rank
        self.world_size = world_size
        assert not use_horovod            # loss.py:478
        self.use_horovod = use_horovod
        self.bidir = bidir                # ring direction of the reference; the gathered sweep covers both
        self.group = None
        self.prev_num_logits = 0
        self.labels = {}

    def get_ground_truth(self, device, dtype, num_logits, negative_only=False) -> torch.Tensor:
        # loss.py:498-502 -- materialising utility, NOT used by forward()
        labels = -torch.ones((num_logits, num_logits), device=device, dtype=dtype)
        if not negative_only:
            labels = 2 * torch.eye(num_logits, device=device, dtype=dtype) + labels
        return labels

    def get_logits(self, image_features, text_features, logit_scale, logit_bias=None):
        # loss.py:504-508 -- materialising utility, NOT used by forward()
        logits = logit_scale * image_features @ text_features.T
        if logit_bias is not None:
            logits = logits + logit_bias
        return logits

    def forward(self, image_features, text_features, logit_scale, logit_bias, output_dict=False):
        if image_features.shape != text_features.shape:
            raise RuntimeError(
                f"image_features {tuple(image_features.shape)} and text_features "
                f"{tuple(text_features.shape)} must have the same shape")
        dev = image_features.device
        if not torch.is_tensor(logit_scale):
            logit_scale = torch.tensor(float(logit_scale), device=dev)
        if logit_bias is not None and not torch.is_tensor(logit_bias):
            logit_bias = torch.tensor(float(logit_bias), device=dev)
        if torch.is_autocast_enabled():
            cdt = torch.get_autocast_dtype("cuda")
        else:
            cdt = torch.promote_types(image_features.dtype, text_features.dtype)
        if not _lib.siglip_supported(cdt, image_features.shape[1]):
            raise RuntimeError(
                f"latteclip_b200.SigLipLoss takes 16-bit features with 8 <= dim <= 768, dim % 8 == 0 "
                f"(got {cdt}, dim {image_features.shape[1]}); run the towers under amp / amp_bf16")
        with torch.autocast(device_type="cuda", enabled=False):
            loss = _FusedSigLip.apply(image_features.to(cdt), text_features.to(cdt), logit_scale,
                                      logit_bias, self.rank, self.world_size, self.group)
        return {"contrastive_loss": loss} if output_dict else loss
