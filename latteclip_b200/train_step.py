"""The step logic of LatteCLIP's ``train_one_epoch_v2`` made DDP-safe, and the ``--accum-freq > 1``
feature-cache path.

Reference (/root/reference/src/training/train.py):
  * ``unwrap_model``                       -- :62-66
  * the per-batch body of the v2 loop       -- :384-530 (``accum_freq == 1``)
  * the feature-cache path of upstream      -- :972-1024 (kept in the file behind
    ``raise NotImplemented()`` at :182, :532, :971; restated in oracle/accum.py)

What "DDP-safe" means here.  The reference reaches ``model.memory_bank``, ``model.tokenizer``,
``model.encode_image`` and ``model.encode_text`` on what ``main.py:318-328`` has wrapped in
``DistributedDataParallel`` (train.py:349, :386, :404, :433): those attributes do not exist on the
wrapper.  ``latteclip_step`` goes through ``unwrap_model`` for them.  Calling ``encode_*`` on the
inner module bypasses DDP's reducer, so the parameter gradients are averaged over the ranks by
``sync_gradients`` (one flat all-reduce) after the backward, and the memory bank stays identical on
every rank because ``prototypes.update_bank`` all-reduces the per-class sums and counts.

Only host logic lives here; every tensor operation of the loss head runs in the CUDA extension
through ``latteclip_b200.prototypes`` / ``latteclip_b200.loss`` (no PyTorch fallback).
"""

from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch

try:
    import torch.distributed as dist
except ImportError:  # pragma: no cover
    dist = None

from . import _lib
from . import prototypes as P
from .loss import ClipLoss
from .zero_shot import zeroshot_class_ids


def unwrap_model(model):
    """train.py:62-66."""
    return model.module if hasattr(model, "module") else model


def backward(total_loss, scaler):
    """train.py:69-73."""
    if scaler is not None:
        scaler.scale(total_loss).backward()
    else:
        total_loss.backward()


@torch.no_grad()
def sync_gradients(parameters, world_size: int, group=None):
    """Average the gradients of ``parameters`` over the ranks with ONE all-reduce of a flat buffer
    (what DDP's reducer would have done had the towers been called through the wrapper).
    Parameters without a gradient on this rank contribute zeros, so every rank reduces the same
    layout."""
    if world_size <= 1:
        return
    params = [p for p in parameters if p.requires_grad]
    if not params:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float()
                      for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= world_size
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n


class ClassTextCache:
    """Tokens of template 0 of every class, tokenised once (train.py:423-438 tokenises the label
    text of every SAMPLE in every step: 2B strings; the distinct strings are the C class names)."""

    def __init__(self, tokenizer: Callable, class_names: Sequence[str], templates: Sequence[Callable],
                 device=None):
        toks = tokenizer([templates[0](c) for c in class_names])
        self.tokens = toks.to(device=device, non_blocking=True) if device is not None else toks
        self.class_names = list(class_names)


def latteclip_step(model, batch, loss: Callable, args, class_names: Sequence[str],
                   templates: Sequence[Callable], proto_snapshot: torch.Tensor,
                   scaler=None, class_text_cache: Optional[ClassTextCache] = None,
                   label_weight_axis: str = "quirk", group=None,
                   do_backward: bool = True) -> Dict[str, torch.Tensor]:
    """One iteration of ``train_one_epoch_v2`` between ``optimizer.zero_grad()`` (train.py:393) and
    the optimizer step (:534): towers -> pseudo-labels -> weights -> mixture + EMA -> two ClipLoss
    calls -> backward (:506) -> memory-bank update (:508-530).

    ``batch`` is the reference's 10-tuple (train.py:372-374); ``proto_snapshot`` the epoch-start
    prototypes (:347-350, ``prototypes.stack_bank``).  The class-name text features are encoded
    once per class and gathered by ``preds`` / ``zs`` inside the mixer kernel (identical values to
    the reference's per-sample re-encodes, :433-438).  Returns the reference's loss dict plus
    ``preds`` and ``zs``.  With ``args.world_size > 1`` the bank update and the tower gradients are
    reduced over ``group``."""
    device = torch.device(args.device)
    inner = unwrap_model(model)                                                # :349,:386,:404,:433
    world = int(getattr(args, "world_size", 1))
    (images, _distill, _texts, _common, _raws, _unused, per_image_texts, per_image_group_texts,
     _meta, zeroshot_classnames) = batch
    images = images.to(device=device, non_blocking=True)
    per_image_texts = per_image_texts.to(device=device, non_blocking=True)
    per_image_group_texts = per_image_group_texts.to(device=device, non_blocking=True)
    dim = per_image_texts.shape[-1]

    bank = P.stack_bank(inner.memory_bank, class_names).to(device)            # :384-387
    if class_text_cache is None:
        class_text_cache = ClassTextCache(inner.tokenizer, class_names, templates, device)

    image_features = inner.encode_image(images, normalize=True)               # :404
    logit_scale = inner.logit_scale.exp()                                      # :405
    class_text = inner.encode_text(class_text_cache.tokens, normalize=True)   # :433-438, once per class
    per_image = inner.encode_text(per_image_texts.reshape(-1, dim), normalize=True)         # :441
    per_group = inner.encode_text(per_image_group_texts.reshape(-1, dim), normalize=True)   # :442
    zs = zeroshot_class_ids(zeroshot_classnames, class_names, device)          # :412-417

    out = P.prototype_step(
        image_features, logit_scale, bank, proto_snapshot, zs, class_text, per_image, per_group, loss,
        alpha=args.alpha, use_image_caption=args.use_image_caption,
        use_batch_caption=args.use_batch_caption, use_template_caption=args.use_template_caption,
        use_zeroshot_pseudolabel=args.use_zeroshot_pseudolabel,
        use_finetune_pseudolabel=args.use_finetune_pseudolabel,
        label_weight_axis=label_weight_axis)                                   # :410-504
    if do_backward:
        backward(out["loss"], scaler)                                          # :506
        sync_gradients(inner.parameters(), world, group)
    P.update_bank(bank, out["preds"], zs, out["t_ft"].detach(), out["t_zs"].detach(),
                              group=group, world_size=world)                   # :508-530
    P.unstack_bank(bank, inner.memory_bank, class_names, inplace=True)     # no host sync
    out["zs"] = zs
    return out


# --------------------------------------------------------------------------------------------
# --accum-freq > 1: cached features as negatives (train.py:972-1024)
# --------------------------------------------------------------------------------------------
_FUSED_DTYPES = [torch.float32, torch.bfloat16, torch.float16]     # what the kernels take


class _AccumClipLoss(torch.autograd.Function):
    """ClipLoss over ``cat(cached[:j] + [live] + cached[j+1:])`` (train.py:1009-1017) where only
    block j carries gradient.  Forward: one sweep over the whole accumulated batch.  Backward: the
    feature gradients exist for the m live rows only, so when ``logit_scale`` needs no gradient the
    recompute covers just the live row block and the live column block (8 m N D executed instead of
    8 N^2 D); ``d logit_scale`` is a sum over the whole matrix, so with it the full backward runs
    and the live rows are sliced out."""

    @staticmethod
    def forward(ctx, live_img, live_txt, logit_scale, work_img, work_txt, j, m):
        lo, hi = j * m, (j + 1) * m
        stats = _lib.clip_fwd(work_img, work_txt, work_img, work_txt, 0, logit_scale, with_nll=True,
                              with_stats=True)
        row_lse, col_lse, loss, row_nll, col_nll, lse_stats = stats
        ctx.save_for_backward(work_img, work_txt, logit_scale.detach(), row_lse, col_lse, row_nll,
                              col_nll, lse_stats)
        ctx.block = (lo, hi)
        ctx.meta = (live_img.dtype, live_txt.dtype, logit_scale.dtype, logit_scale.shape)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        (work_img, work_txt, scale, row_lse, col_lse, row_nll, col_nll, lse_stats) = ctx.saved_tensors
        # (autograd's version check on the saved work buffers raises if another micro_loss ran
        # before this backward: the reference calls backward inside the loop, train.py:1023)
        lo, hi = ctx.block
        n_all = work_img.shape[0]
        di_t, dt_t, s_dtype, s_shape = ctx.meta
        gdt = di_t if di_t == dt_t and di_t in (torch.float32, torch.bfloat16, torch.float16) else torch.float32
        if ctx.needs_input_grad[2]:
            d_img, d_txt, d_scale = _lib.clip_bwd(
                work_img, work_txt, work_img, work_txt, 0, scale, row_lse, col_lse, grad_out, 1.0, True,
                grad_dtype=gdt, row_nll_all=row_nll, col_nll_all=col_nll, lse_stats=lse_stats)
            d_img, d_txt = d_img[lo:hi], d_txt[lo:hi]
            d_scale = d_scale.reshape(s_shape).to(s_dtype)
        else:
            # rows and columns of the live block only; the block loss the kernels differentiate is a
            # mean over m rows, the accumulated loss a mean over N: grad_mult = m / N
            d_img, d_txt, _ = _lib.clip_bwd(
                work_img[lo:hi], work_txt[lo:hi], work_img, work_txt, lo, scale, row_lse, col_lse,
                grad_out, (hi - lo) / n_all, True, grad_dtype=gdt, row_nll_all=row_nll,
                col_nll_all=col_nll)
            d_scale = None
        need = ctx.needs_input_grad
        return (d_img.to(di_t) if need[0] else None, d_txt.to(dt_t) if need[1] else None, d_scale,
                None, None, None, None)


class FeatureAccumulator:
    """The ``--accum-freq`` feature cache of train.py:972-1024 for ``ClipLoss``-style losses.

        acc = FeatureAccumulator(loss, accum_freq)
        for images, texts in micro_batches:              # train.py:974-987
            with torch.no_grad():
                acc.cache(model(images, texts))
        for j, (images, texts) in enumerate(micro_batches):     # :998-1023
            losses = acc.micro_loss(j, model(images, texts))
            backward(losses["loss"], scaler)
        acc.reset()                                      # :1046-1048

    ``model_out`` is the reference's dict (``image_features``, ``text_features``, ``logit_scale``
    and optionally ``logit_bias``).  With a single-process ``latteclip_b200.ClipLoss`` the cached
    features live in one [accum_freq * m, D] buffer per tower and ``micro_loss`` only rewrites the
    live block; otherwise (multi-rank, SigLIP, any other callable) the features are concatenated as
    in the reference and handed to ``loss``."""

    def __init__(self, loss: Callable, accum_freq: int):
        self.loss = loss
        self.accum_freq = int(accum_freq)
        self.reset()

    def reset(self):
        self.features: Dict[str, List[torch.Tensor]] = {}
        self._work = None
        self._restored = None

    def cache(self, model_out: Dict[str, torch.Tensor]):
        """train.py:975-987: keep the no-grad features of one micro-batch."""
        for key, val in model_out.items():
            if key in ("logit_scale", "logit_bias"):
                continue
            self.features.setdefault(key, []).append(val.detach())

    def ready(self) -> bool:
        n = len(next(iter(self.features.values()), []))
        return n == self.accum_freq

    def _fused(self) -> bool:
        lf = self.loss
        if not isinstance(lf, ClipLoss) or lf.world_size > 1 or lf.normalize_features:
            return False
        if set(self.features) != {"image_features", "text_features"}:
            return False
        fi, ft = self.features["image_features"], self.features["text_features"]
        shapes = {tuple(x.shape) for x in fi + ft}
        dtypes = {x.dtype for x in fi + ft}
        return len(shapes) == 1 and len(dtypes) == 1 and fi[0].dtype in _FUSED_DTYPES \
            and not torch.is_autocast_enabled()

    def micro_loss(self, j: int, model_out: Dict[str, torch.Tensor], output_dict: bool = True):
        """train.py:1002-1019: the loss of micro-batch j with the other micro-batches' cached
        features as negatives.  Returns the loss dict with ``"loss"`` = the sum (:1018-1019)."""
        model_out = dict(model_out)
        no_accum = {"logit_scale": model_out.pop("logit_scale")}
        if "logit_bias" in model_out:
            no_accum["logit_bias"] = model_out.pop("logit_bias")
        if self._fused():
            live_i, live_t = model_out["image_features"], model_out["text_features"]
            m = live_i.shape[0]
            if self._work is None:
                self._work = (torch.cat(self.features["image_features"]).contiguous(),
                              torch.cat(self.features["text_features"]).contiguous())
                self._restored = None
            work_i, work_t = self._work
            if self._restored is not None and self._restored != j:
                k = self._restored      # put the cached block of the previous micro-batch back
                work_i[k * m:(k + 1) * m].copy_(self.features["image_features"][k])
                work_t[k * m:(k + 1) * m].copy_(self.features["text_features"][k])
            work_i[j * m:(j + 1) * m].copy_(live_i.detach())
            work_t[j * m:(j + 1) * m].copy_(live_t.detach())
            self._restored = j
            scale = no_accum["logit_scale"]
            if not torch.is_tensor(scale):
                scale = torch.tensor(float(scale), device=live_i.device)
            total = _AccumClipLoss.apply(live_i, live_t, scale, work_i, work_t, j, m)
            losses = {"contrastive_loss": total}
        else:
            inputs = {}
            for key, accumulated in self.features.items():
                inputs[key] = torch.cat(accumulated[:j] + [model_out[key]] + accumulated[j + 1:])
            losses = self.loss(**inputs, **no_accum, output_dict=True)
        total_loss = sum(losses.values())
        losses["loss"] = total_loss
        return losses if output_dict else total_loss
