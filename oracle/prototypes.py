"""Oracle (TEST INFRASTRUCTURE ONLY): CPU restatement of LatteCLIP's prototype path.

The reference has no function boundary here: the code is inline in
``train_one_epoch_v2`` (/root/reference/src/training/train.py:306-636) plus the helper
``compute_text_weights`` (train.py:292-303).  This module restates it over tensors and
integer index vectors (class names -> ids), keeping the reference's operation order so
that fp32 results track the verbatim expressions closely.

Two behaviours of the reference are reproduced, not fixed:
  * quirk 1 (train.py:476, :481): ``label_text_weight * label_text_features``
    multiplies a [B] vector with a [B, D] matrix WITHOUT unsqueeze -> torch broadcasts
    along the LAST axis (needs B == D).  ``label_weight_axis="quirk"`` is that verbatim
    behaviour; ``"row"`` is the evident intent (``w[:, None]``) and the only mode
    usable when B != D (where the reference expression raises).
  * quirk 2 (train.py:481 vs :473): the zero-shot mixture uses ``label_text_weight``
    (the fine-tune weight) in its numerator but ``label_text_weight_zeroshot`` in its
    denominator.
"""

from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .clip_loss import clip_loss_reference


def build_classifier(bank: torch.Tensor) -> torch.Tensor:
    """train.py:384-389 (also zero_shot.py:138-145): rows of the stacked memory bank,
    L2-normalised along dim=1.  Returns P_hat [C, D] (the reference then uses P_hat.T)."""
    return F.normalize(bank, dim=1)


def pseudo_label(image_features: torch.Tensor, classifier: torch.Tensor,
                 scale: float = 100.0) -> torch.Tensor:
    """train.py:410-411: ``logits = 100.0 * image_features @ classifier; argmax(dim=1)``.
    ``classifier`` here is P_hat [C, D]; first maximal index wins (torch.argmax)."""
    logits = scale * image_features @ classifier.T
    return logits.argmax(dim=1)


def text_margins(text_features: torch.Tensor, prototypes: torch.Tensor) -> torch.Tensor:
    """compute_text_weights (train.py:292-303): top-1 minus top-2 of
    ``text_features @ prototypes.T`` per row.  The reference evaluates the product as a
    batched 1xDxC bmm and ignores its ``preds`` argument (train.py:301-303)."""
    sim = text_features @ prototypes.T
    top2 = torch.topk(sim, 2, dim=1).values
    return top2[:, 0] - top2[:, 1]


def mix_and_ema(label_ft: torch.Tensor, label_zs: torch.Tensor,
                per_image: torch.Tensor, per_group: torch.Tensor,
                w_lbl: torch.Tensor, w_lbl_zs: torch.Tensor,
                w_img: torch.Tensor, w_grp: torch.Tensor,
                bank_ft: torch.Tensor, bank_zs: torch.Tensor,
                alpha: float, label_weight_axis: str = "row"):
    """train.py:472-488.

    total      = w_lbl    + w_img + w_grp                                   (:472)
    total_zs   = w_lbl_zs + w_img + w_grp                                   (:473)
    mix_ft     = (w_lbl (*) L_ft + P * w_img[:,None] + G * w_grp[:,None]) / total[:,None]      (:476-479)
    mix_zs     = (w_lbl (*) L_zs + P * w_img[:,None] + G * w_grp[:,None]) / total_zs[:,None]   (:481-484)
    T_ft       = M_ft + alpha * (mix_ft - M_ft)                             (:487)
    T_zs       = M_zs + alpha * (mix_zs - M_zs)                             (:488)

    (*) is the quirk-1 broadcast: "quirk" -> along the last axis, "row" -> per row.
    Note w_lbl (not w_lbl_zs) in the zero-shot numerator (quirk 2).
    """
    if label_weight_axis == "quirk":
        if label_ft.shape[0] != label_ft.shape[1]:
            raise RuntimeError("quirk broadcast needs B == D (train.py:476)")
        wl = w_lbl  # [B] broadcast against [B, D] -> multiplies column d by w_lbl[d]
    elif label_weight_axis == "row":
        wl = w_lbl[:, None]
    else:
        raise ValueError(label_weight_axis)
    total = w_lbl + w_img + w_grp
    total_zs = w_lbl_zs + w_img + w_grp
    mix_ft = wl * label_ft + per_image * w_img.unsqueeze(1) + per_group * w_grp.unsqueeze(1)
    mix_ft = mix_ft / total.unsqueeze(1)
    mix_zs = wl * label_zs + per_image * w_img.unsqueeze(1) + per_group * w_grp.unsqueeze(1)
    mix_zs = mix_zs / total_zs.unsqueeze(1)
    t_ft = bank_ft + alpha * (mix_ft - bank_ft)
    t_zs = bank_zs + alpha * (mix_zs - bank_zs)
    return t_ft, t_zs


def update_bank(bank: torch.Tensor, preds: torch.Tensor, zs: torch.Tensor,
                t_ft: torch.Tensor, t_zs: torch.Tensor) -> torch.Tensor:
    """train.py:508-530 (under no_grad).  For every class c touched by this batch:
    bank[c] = normalize( (sum_{i: zs_i=c} T_zs[i] + sum_{i: preds_i=c} T_ft[i]) / count_c ).
    Untouched classes keep their row.  The per-sample accumulation order of the
    reference loop is kept (for each i: the zero-shot row first, then the pseudo-label
    row, train.py:524-525)."""
    C, D = bank.shape
    sums = torch.zeros(C, D, dtype=t_ft.dtype)
    cnt = torch.zeros(C, dtype=torch.long)
    preds_l = preds.tolist()
    zs_l = zs.tolist()
    for i in range(len(preds_l)):
        sums[zs_l[i]] += t_zs[i]
        sums[preds_l[i]] += t_ft[i]
        cnt[zs_l[i]] += 1
        cnt[preds_l[i]] += 1
    out = bank.detach().clone()
    for c in range(C):
        if cnt[c] > 0:
            out[c] = F.normalize(sums[c] / cnt[c].item(), dim=0).to(out.dtype)
    return out


def update_bank_vectorised(bank, preds, zs, t_ft, t_zs):
    """Same result as ``update_bank`` up to fp32 summation order (index_add_); used
    for sizes where the Python loop is too slow."""
    C, D = bank.shape
    sums = torch.zeros(C, D, dtype=t_ft.dtype)
    sums.index_add_(0, zs, t_zs)
    sums.index_add_(0, preds, t_ft)
    cnt = torch.bincount(zs, minlength=C) + torch.bincount(preds, minlength=C)
    out = bank.detach().clone()
    touched = cnt > 0
    mean = sums[touched] / cnt[touched].to(sums.dtype)[:, None]
    out[touched] = F.normalize(mean, dim=1).to(out.dtype)
    return out


def prototype_step(image_features: torch.Tensor,
                   logit_scale: torch.Tensor,
                   bank: torch.Tensor,
                   proto_snapshot: torch.Tensor,
                   zs: torch.Tensor,
                   class_text: torch.Tensor,
                   per_image: torch.Tensor,
                   per_group: torch.Tensor,
                   alpha: float = 0.01,
                   use_image_caption: float = 1.0,
                   use_batch_caption: float = 1.0,
                   use_template_caption: float = 1.0,
                   use_zeroshot_pseudolabel: float = 1.0,
                   use_finetune_pseudolabel: float = 1.0,
                   label_weight_axis: str = "row",
                   update: bool = True,
                   preds: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """One step of the hot path of train_one_epoch_v2 (train.py:384-530), SURVEY.md
    appendix A steps 1-9, over tensors.

    image_features [B,D]  = model.encode_image(images, normalize=True)      (:404)
    logit_scale    0-dim  = model.logit_scale.exp()                          (:405)
    bank           [C,D]  current memory bank rows                           (:384-387)
    proto_snapshot [C,D]  bank as of epoch start                             (:347-350)
    zs             [B]    class id of the dataloader's frozen-CLIP top-1     (:416-417)
    class_text     [C,D]  normalize(encode_text(template0(class c)))         (:423-438)
    per_image      [B,D]  LMM image-description text features                (:441)
    per_group      [B,D]  LMM group-description text features                (:442)
    """
    classifier = build_classifier(bank.detach())                    # :384-389
    if preds is None:
        preds = pseudo_label(image_features.detach(), classifier)   # :410-411
    label_ft = class_text[preds]                                     # :420-438
    label_zs = class_text[zs]
    bank_ft = bank[preds]                                            # :428-431
    bank_zs = bank[zs]

    w_img = (text_margins(per_image, proto_snapshot).detach() + 1e-6) * use_image_caption      # :444,:463
    w_grp = (text_margins(per_group, proto_snapshot).detach() + 1e-6) * use_batch_caption      # :445,:460
    w_lbl = (text_margins(label_ft, proto_snapshot).detach() + 1e-6) * use_template_caption    # :448,:468
    w_lbl_zs = (text_margins(label_zs, proto_snapshot).detach() + 1e-6) * use_template_caption # :449,:469

    t_ft, t_zs = mix_and_ema(label_ft, label_zs, per_image, per_group,
                             w_lbl, w_lbl_zs, w_img, w_grp, bank_ft, bank_zs,
                             alpha, label_weight_axis)               # :472-488

    loss_ft = clip_loss_reference(image_features, t_ft, logit_scale)   # :491-494
    loss_zs = clip_loss_reference(image_features, t_zs, logit_scale)   # :496-499
    zeroshot = loss_zs * use_zeroshot_pseudolabel                      # :501
    total = (loss_ft + zeroshot) * use_finetune_pseudolabel            # :502

    out = dict(preds=preds, w_img=w_img, w_grp=w_grp, w_lbl=w_lbl, w_lbl_zs=w_lbl_zs,
               t_ft=t_ft, t_zs=t_zs, contrastive_loss=loss_ft, zeroshot=zeroshot,
               loss=total)
    if update:
        fn = update_bank if image_features.shape[0] <= 8192 else update_bank_vectorised
        out["bank"] = fn(bank.detach(), preds, zs, t_ft.detach(), t_zs.detach())  # :508-530
    return out
