"""LatteCLIP's prototype / pseudo-label / mixture / EMA / memory-bank path as functions.

The reference has no function boundary here: the code is inline in
``train_one_epoch_v2`` (/root/reference/src/training/train.py:306-636) plus the helper
``compute_text_weights`` (train.py:292-303).  This module defines the boundary
(boundary #2).  Each function states the reference lines it replaces;
all arithmetic runs in the CUDA extension through ``latteclip_b200._lib`` -- there is no
PyTorch fallback.

The memory bank is handled as one stacked fp32 tensor ``bank[C, D]`` in class order;
``stack_bank`` / ``unstack_bank`` convert from / to the reference's
``model.memory_bank`` ParameterDict (model.py:489-499) so checkpoints keep their format.
"""

from __future__ import annotations

from typing import Dict, Iterable, Optional

import torch
import torch.nn as nn

try:
    import torch.distributed as dist
except ImportError:  # pragma: no cover
    dist = None

from . import _lib


# ------------------------------------------------------------------------------ bank storage
def stack_bank(memory_bank, class_names: Iterable[str]) -> torch.Tensor:
    """train.py:347-350 / :384-387: stack the per-class Parameters in class order."""
    return torch.stack([memory_bank[c].detach() for c in class_names]).to(torch.float32).contiguous()


def unstack_bank(bank: torch.Tensor, memory_bank, class_names: Iterable[str], touched=None,
                 inplace: bool = False):
    """Write rows of the stacked bank back into the ParameterDict (train.py:529-530 assigns a
    new tensor per touched class).  ``touched``: optional bool/float [C] mask (counts > 0); reading
    it costs a device-to-host sync.  ``inplace=True`` copies EVERY row into the existing Parameters
    with one foreach copy instead (no sync, no allocation): rows of untouched classes are
    bit-identical in the stacked bank, so the stored values are the same as the reference's."""
    if inplace:
        names = list(class_names)
        dst = [memory_bank[c].data for c in names]
        torch._foreach_copy_(dst, [r.to(d.dtype) for r, d in zip(bank.unbind(0), dst)])
        return memory_bank
    mask = None if touched is None else (touched > 0).tolist()
    for k, c in enumerate(class_names):
        if mask is None or mask[k]:
            memory_bank[c] = nn.Parameter(bank[k].clone())
    return memory_bank


# ------------------------------------------------------------------------------ kernels
# 16-bit operand planes of the epoch-start prototype snapshot (train.py:347-350): split once per
# snapshot tensor, not once per call.  The planes ride on the tensor object itself (with its
# in-place version counter), so they die with it and a new tensor at a recycled address never
# sees stale planes.
def _snapshot_planes(proto_snapshot: torch.Tensor):
    cached = getattr(proto_snapshot, "_latte_planes", None)
    if cached is not None and cached[0] == proto_snapshot._version and \
            cached[1] == proto_snapshot.data_ptr():
        return cached[2]
    pl = _lib.nxc_split_prototypes(proto_snapshot, normalize=False)
    try:
        proto_snapshot._latte_planes = (proto_snapshot._version, proto_snapshot.data_ptr(), pl)
    except AttributeError:      # exotic tensor subclass without a __dict__: split per call
        pass
    return pl


def step_similarities(image_features, bank, proto_snapshot, per_image, per_group, class_text):
    """The step's five N x C products (train.py:410-411 and the distinct compute_text_weights calls
    of :444-449) -> (preds int64 [B], margin(per_image) [B], margin(per_group) [B],
    margin(class_text) [C]).  With C <= 64 they run as ONE launch per operand class
    (latte_nxc_multi: stacked jobs, prototypes split once, rows streamed once); otherwise one
    launch per product."""
    xs = (image_features, per_image, per_group, class_text)
    c = bank.shape[0]
    if all(_lib.nxc_multi_supported(x.detach(), c) for x in xs) and proto_snapshot.shape == bank.shape:
        cls_planes = _lib.nxc_split_prototypes(bank, normalize=True)          # :384-389 fused
        snap_planes = _snapshot_planes(proto_snapshot)
        jobs = [dict(x=image_features, planes=cls_planes, scale=100.0, argmax=True),
                dict(x=per_image, planes=snap_planes, margin=True),
                dict(x=per_group, planes=snap_planes, margin=True),
                dict(x=class_text, planes=snap_planes, margin=True)]
        res = [None] * 4
        for direct in (True, False):      # 16-bit rows are MMA operands as they are; fp32 rows are converted
            idx = [k for k, jb in enumerate(jobs) if (jb["x"].dtype != torch.float32) == direct]
            if idx:
                for k, out in zip(idx, _lib.nxc_multi([jobs[k] for k in idx])):
                    res[k] = out
        return res[0][0], res[1][1], res[2][1], res[3][1]
    classifier = build_classifier(bank)                                         # :384-389
    preds = pseudo_label(image_features, classifier, 100.0)                     # :410-411
    return (preds, text_margins(per_image, proto_snapshot), text_margins(per_group, proto_snapshot),
            text_margins(class_text, proto_snapshot))


def build_classifier(bank: torch.Tensor) -> torch.Tensor:
    """train.py:384-389 (and zero_shot.py:138-145): ``F.normalize(stack(bank), dim=1)``.
    Returns P_hat [C, D] fp32 (the reference then uses ``P_hat.T``)."""
    return _lib.normalize_rows(bank)


def pseudo_label(image_features: torch.Tensor, classifier: torch.Tensor,
                 scale: float = 100.0) -> torch.Tensor:
    """train.py:410-411: ``(100.0 * image_features @ classifier).argmax(dim=1)`` with
    ``classifier = P_hat.T``; here ``classifier`` is P_hat [C, D].  int64 [B]; the first
    maximal index wins.  The [B, C] logits are never stored."""
    am, _, _ = _lib.nxc_argmax_margin(image_features, classifier, scale=scale,
                                      want_argmax=True, want_margin=False)
    return am


def text_margins(text_features: torch.Tensor, prototypes: torch.Tensor,
                 row_index: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``compute_text_weights`` (train.py:292-303): top-1 minus top-2 of
    ``text_features @ prototypes.T`` per row (its ``preds`` argument is unused by the
    reference, train.py:301-303).  With ``row_index`` the rows are gathered first:
    ``text_features[row_index]`` (class-name text features by label, train.py:420-438)."""
    _, mg, _ = _lib.nxc_argmax_margin(text_features, prototypes, scale=1.0, row_index=row_index,
                                      want_argmax=False, want_margin=True)
    return mg


def zero_shot_topk(image_features: torch.Tensor, classifier: torch.Tensor, k: int,
                   scale: float = 100.0):
    """zero_shot.py:40 + :14-20 / train.py:1352-1358: ``(100 * I @ classifier).topk(k)``
    without storing the logits.  Returns (idx int64 [B, k], val fp32 [B, k])."""
    return _lib.nxc_topk(image_features, classifier, k, scale=scale)


class _MixEma(torch.autograd.Function):
    @staticmethod
    def forward(ctx, class_text, per_image, per_group, bank, preds, zs, w_lbl, w_lbl_zs, w_img,
                w_grp, alpha, label_weight_axis):
        t_ft, t_zs = _lib.mix_ema_fwd(class_text, per_image, per_group, bank, preds, zs, w_lbl,
                                      w_lbl_zs, w_img, w_grp, alpha, label_weight_axis)
        ctx.save_for_backward(preds, zs, w_lbl, w_lbl_zs, w_img, w_grp)
        ctx.meta = (alpha, label_weight_axis, class_text.shape[0], class_text.dtype,
                    bank.dtype)
        return t_ft, t_zs

    @staticmethod
    def backward(ctx, d_t_ft, d_t_zs):
        preds, zs, w_lbl, w_lbl_zs, w_img, w_grp = ctx.saved_tensors
        alpha, axis, num_classes, ct_dtype, bank_dtype = ctx.meta
        want_bank = ctx.needs_input_grad[3]
        d_ct, d_pi, d_pg, d_bank = _lib.mix_ema_bwd(d_t_ft, d_t_zs, preds, zs, w_lbl, w_lbl_zs,
                                                    w_img, w_grp, alpha, axis, num_classes,
                                                    want_bank=want_bank)
        return (d_ct.to(ct_dtype) if ctx.needs_input_grad[0] else None,
                d_pi if ctx.needs_input_grad[1] else None,
                d_pg if ctx.needs_input_grad[2] else None,
                d_bank.to(bank_dtype) if want_bank else None,
                None, None, None, None, None, None, None, None)


def mix_and_ema(class_text: torch.Tensor, per_image: torch.Tensor, per_group: torch.Tensor,
                bank: torch.Tensor, preds: torch.Tensor, zs: torch.Tensor,
                w_lbl: torch.Tensor, w_lbl_zs: torch.Tensor, w_img: torch.Tensor,
                w_grp: torch.Tensor, alpha: float, label_weight_axis: str = "row"):
    """train.py:472-488 with the gathers of :420-431 fused:

        L_ft = class_text[preds]; L_zs = class_text[zs]; M_ft = bank[preds]; M_zs = bank[zs]
        T_ft = M_ft + alpha * ((w_lbl (*) L_ft + w_img*P + w_grp*G) / (w_lbl    + w_img + w_grp) - M_ft)
        T_zs = M_zs + alpha * ((w_lbl (*) L_zs + w_img*P + w_grp*G) / (w_lbl_zs + w_img + w_grp) - M_zs)

    ``label_weight_axis="quirk"`` is the reference's literal broadcast (needs B == D,
    train.py:476); ``"row"`` is ``w_lbl[:, None]``.  Differentiable w.r.t. class_text,
    per_image, per_group (and bank, if it requires grad); the weights are detached in the
    reference (train.py:444-449)."""
    if label_weight_axis not in ("row", "quirk"):
        raise ValueError(label_weight_axis)
    if label_weight_axis == "quirk" and per_image.shape[0] != per_image.shape[1]:
        # same failure the reference expression produces when B != D
        raise RuntimeError(
            f"The size of tensor a ({per_image.shape[0]}) must match the size of tensor b "
            f"({per_image.shape[1]}) at non-singleton dimension 1")
    return _MixEma.apply(class_text, per_image, per_group, bank, preds, zs,
                         w_lbl.detach(), w_lbl_zs.detach(), w_img.detach(), w_grp.detach(),
                         float(alpha), label_weight_axis)


@torch.no_grad()
def update_bank(bank: torch.Tensor, preds: torch.Tensor, zs: torch.Tensor,
                t_ft: torch.Tensor, t_zs: torch.Tensor, group=None, world_size: int = 1):
    """train.py:508-530: for every class c touched by the batch
    ``bank[c] = normalize((sum_{zs_i=c} T_zs[i] + sum_{preds_i=c} T_ft[i]) / count_c)``; other
    rows unchanged.  ``bank`` (fp32 [C, D], contiguous) is updated in place; returns
    (bank, counts).  With world_size > 1 the per-class sums and counts are all-reduced first,
    so every rank ends with the bank a single process would compute on the concatenated batch
    (the reference loop is single-process only)."""
    sums, counts = _lib.bank_accumulate(t_ft, t_zs, preds, zs, bank.shape[0])
    if world_size > 1:
        packed = torch.cat([sums, counts[:, None]], dim=1)
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        sums, counts = packed[:, :-1].contiguous(), packed[:, -1].contiguous()
    _lib.bank_finalize(sums, counts, bank)
    return bank, counts


def prototype_step(image_features: torch.Tensor,
                   logit_scale: torch.Tensor,
                   bank: torch.Tensor,
                   proto_snapshot: torch.Tensor,
                   zs: torch.Tensor,
                   class_text: torch.Tensor,
                   per_image: torch.Tensor,
                   per_group: torch.Tensor,
                   loss_fn,
                   alpha: float = 0.01,
                   use_image_caption: float = 1.0,
                   use_batch_caption: float = 1.0,
                   use_template_caption: float = 1.0,
                   use_zeroshot_pseudolabel: float = 1.0,
                   use_finetune_pseudolabel: float = 1.0,
                   label_weight_axis: str = "row") -> Dict[str, torch.Tensor]:
    """The hot path of one ``train_one_epoch_v2`` iteration (train.py:384-504), steps 1-8 of
    SURVEY.md appendix A.  ``class_text[c]`` is the text feature of class c's label template
    (the reference re-encodes it per sample, train.py:433-438; the values are identical).
    Returns the loss dict of the reference (keys ``contrastive_loss``, ``zeroshot``, ``loss``,
    train.py:491-504) plus ``preds``, ``t_ft``, ``t_zs`` and the weights.  Call
    ``out["loss"].backward()`` and then ``update_bank`` (train.py:506-530).

    Feature dtype: ``t_ft`` / ``t_zs`` come back in the dtype of ``per_image``.  In the reference's amp
    flow the text features are fp32 (``F.normalize`` runs in fp32 under autocast, model.py:420-437), so
    T and the bank update stay fp32 and only the ClipLoss operands are rounded to the autocast dtype
    (inside ``latte_prep_features``).  Feeding 16-bit features (as the benchmarks do) rounds T -- and
    the EMA step of ~alpha -- to that dtype: about 6e-4 relative on the updated bank rows."""
    # pseudo-labels against the normalised bank (:384-389, :410-411) and the margins against the
    # epoch-start snapshot (:347-350), weights detached (:444-449): one stacked launch
    preds, m_img, m_grp, cls_margin = step_similarities(image_features, bank, proto_snapshot, per_image,
                                                        per_group, class_text)
    w_img = (m_img + 1e-6) * use_image_caption                                  # :444,:463
    w_grp = (m_grp + 1e-6) * use_batch_caption                                  # :445,:460
    # cls_margin: one margin per class, gathered below
    w_lbl = (cls_margin[preds] + 1e-6) * use_template_caption                   # :448,:468
    w_lbl_zs = (cls_margin[zs] + 1e-6) * use_template_caption                   # :449,:469
    t_ft, t_zs = mix_and_ema(class_text, per_image, per_group, bank, preds, zs,
                             w_lbl, w_lbl_zs, w_img, w_grp, alpha, label_weight_axis)   # :472-488
    losses = loss_fn(image_features=image_features, text_features=t_ft,
                     logit_scale=logit_scale, output_dict=True)                 # :491-494
    losses_zs = loss_fn(image_features=image_features, text_features=t_zs,
                        logit_scale=logit_scale, output_dict=True)              # :496-499
    losses["zeroshot"] = sum(losses_zs.values()) * use_zeroshot_pseudolabel     # :501
    total = sum(losses.values()) * use_finetune_pseudolabel                     # :502
    losses["loss"] = total                                                      # :504
    losses.update(preds=preds, t_ft=t_ft, t_zs=t_zs, w_img=w_img, w_grp=w_grp,
                  w_lbl=w_lbl, w_lbl_zs=w_lbl_zs)
    return losses


# ------------------------------------------------------------------------------ CUDA-graph step
class _ReplayStep(torch.autograd.Function):
    """Forward = one replay of the captured prototype_step + backward (+ bank update); backward hands
    out the gradients that replay left in the static buffers, scaled by the incoming gradient."""

    @staticmethod
    def forward(ctx, runner, image_features, logit_scale, class_text, per_image, per_group):
        runner._replay(image_features, logit_scale, class_text, per_image, per_group)
        ctx.runner = runner
        ctx.serial = runner.serial
        st = runner.static_out
        return st["loss"].detach().clone(), st["contrastive_loss"].detach().clone(), st["zeroshot"].detach().clone()

    @staticmethod
    def backward(ctx, g_loss, _g_c, _g_z):
        r = ctx.runner
        if ctx.serial != r.serial:
            raise RuntimeError("GraphedPrototypeStep: backward after a later step replaced the static gradients")
        return (None,) + tuple(g_loss * leaf.grad for leaf in r.leaves)


class GraphedPrototypeStep:
    """``prototype_step`` + ``loss.backward()`` (+ ``update_bank``) captured ONCE in a CUDA graph and
    replayed per step: at the reference's fine-tuning shape (batch 512, 47 classes) the step is
    ~40 short kernels and its cost is launch gaps, not arithmetic.

        step = GraphedPrototypeStep(loss_fn, alpha=args.alpha, label_weight_axis="quirk")
        out = step(image_features, logit_scale, bank, proto_snapshot, zs, class_text, per_image, per_group)
        out["loss"].backward()          # hands the gradients computed by the replay to the towers

    The replay already ran the loss head's backward (with upstream gradient 1); the returned
    ``out["loss"]`` is connected to the five differentiable inputs through an autograd node that
    scales those gradients by whatever arrives (a GradScaler factor, a loss weight).  ``bank`` and
    ``proto_snapshot`` are baked into the graph by address (the bank is updated in place, as
    ``update_bank`` does); new shapes, dtypes or tensors re-capture.  Single-process ClipLoss only
    (collectives are not captured)."""

    def __init__(self, loss_fn, alpha: float = 0.01, use_image_caption: float = 1.0,
                 use_batch_caption: float = 1.0, use_template_caption: float = 1.0,
                 use_zeroshot_pseudolabel: float = 1.0, use_finetune_pseudolabel: float = 1.0,
                 label_weight_axis: str = "row", with_bank_update: bool = True):
        if getattr(loss_fn, "world_size", 1) > 1:
            raise NotImplementedError("GraphedPrototypeStep: world_size > 1 is not captured")
        self.loss_fn = loss_fn
        self.kw = dict(alpha=alpha, use_image_caption=use_image_caption, use_batch_caption=use_batch_caption,
                       use_template_caption=use_template_caption,
                       use_zeroshot_pseudolabel=use_zeroshot_pseudolabel,
                       use_finetune_pseudolabel=use_finetune_pseudolabel, label_weight_axis=label_weight_axis)
        self.with_bank_update = with_bank_update
        self.key = None
        self.graph = None
        self.serial = 0

    def _eager(self):
        for x in self.leaves:
            x.grad = None
        out = prototype_step(self.leaves[0], self.leaves[1], self.bank, self.snapshot, self.s_zs,
                             self.leaves[2], self.leaves[3], self.leaves[4], self.loss_fn, **self.kw)
        out["loss"].backward()
        if self.with_bank_update:
            update_bank(self.bank, out["preds"], self.s_zs, out["t_ft"].detach(), out["t_zs"].detach())
        return out

    def _capture(self, image_features, logit_scale, bank, proto_snapshot, zs, class_text, per_image, per_group):
        def leaf(x):
            return x.detach().clone().requires_grad_(True)
        self.leaves = [leaf(image_features), leaf(logit_scale), leaf(class_text), leaf(per_image), leaf(per_group)]
        self.s_zs = zs.detach().clone()
        self.bank, self.snapshot = bank, proto_snapshot
        keep = bank.detach().clone()
        side = torch.cuda.Stream(device=bank.device)
        side.wait_stream(torch.cuda.current_stream(bank.device))
        with torch.cuda.stream(side):
            for _ in range(3):                       # warm-up: workspaces, lazy module state
                self._eager()
        torch.cuda.current_stream(bank.device).wait_stream(side)
        bank.copy_(keep)                             # the warm-up steps must not move the bank
        for x in self.leaves:
            x.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = self._eager()
        bank.copy_(keep)                             # (capture does not execute, but keep it explicit)

    def _replay(self, image_features, logit_scale, class_text, per_image, per_group):
        with torch.no_grad():
            for dst, src in zip(self.leaves, (image_features, logit_scale, class_text, per_image, per_group)):
                dst.copy_(src)
        self.graph.replay()
        self.serial += 1

    def __call__(self, image_features, logit_scale, bank, proto_snapshot, zs, class_text, per_image, per_group):
        key = (tuple(image_features.shape), image_features.dtype, tuple(class_text.shape), class_text.dtype,
               per_image.dtype, per_group.dtype, logit_scale.dtype, tuple(logit_scale.shape), bank.data_ptr(),
               proto_snapshot.data_ptr(), proto_snapshot._version, tuple(zs.shape))
        if key != self.key:
            self._capture(image_features, logit_scale, bank, proto_snapshot, zs, class_text, per_image, per_group)
            self.key = key
        self.s_zs.copy_(zs)
        loss, contrastive, zeroshot = _ReplayStep.apply(self, image_features, logit_scale, class_text,
                                                        per_image, per_group)
        st = self.static_out
        return {"loss": loss, "contrastive_loss": contrastive, "zeroshot": zeroshot, "preds": st["preds"],
                "t_ft": st["t_ft"].detach(), "t_zs": st["t_zs"].detach(), "w_img": st["w_img"],
                "w_grp": st["w_grp"], "w_lbl": st["w_lbl"], "w_lbl_zs": st["w_lbl_zs"]}
