"""Oracle (TEST INFRASTRUCTURE ONLY): CPU restatement of open_clip's ClipLoss.

Follows /root/reference/src/open_clip/loss.py:
  * gather_features            loss.py:19-63
  * ClipLoss.get_ground_truth  loss.py:89-100
  * ClipLoss.get_logits        loss.py:102-118
  * ClipLoss.forward           loss.py:120-130

Plain torch on CPU, any float dtype (fp32 to mirror the reference, fp64 as the
high-precision yardstick).  Multi-rank semantics are emulated inside ONE process by
building, for every rank, exactly the tensors that rank would see after
``gather_features`` (attached / detached slots as in loss.py:48-61) and summing the
per-rank losses before a single ``backward`` -- the leaf gradients are then what the
reference's all_gather backward (a reduce-scatter SUM) delivers to each rank.
SURVEY.md appendix B probe 5 checked this emulation against a real 4-process gloo run
of the reference; tests/golden/make_golden.py records the reference's gloo runs (2 and 4
processes) and tests/test_oracle_golden.py holds this emulation to them.
"""

from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def _cross_entropy_mean(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """F.cross_entropy(logits, labels) with mean reduction, spelled out
    (loss.py:127-128): mean_i( logsumexp_j(logits_ij) - logits_i,label_i )."""
    lse = torch.logsumexp(logits, dim=1)
    picked = logits.gather(1, labels[:, None]).squeeze(1)
    return (lse - picked).mean()


def clip_loss_reference(image_features: torch.Tensor,
                        text_features: torch.Tensor,
                        logit_scale: torch.Tensor) -> torch.Tensor:
    """world_size == 1 branch (loss.py:115-116, 124-129).

    Note the precedence in the reference: ``logit_scale * image_features @ text.T``
    scales the A operand before the GEMM.
    """
    logits_per_image = (logit_scale * image_features) @ text_features.T
    logits_per_text = (logit_scale * text_features) @ image_features.T
    n = logits_per_image.shape[0]
    labels = torch.arange(n, dtype=torch.long)
    return (_cross_entropy_mean(logits_per_image, labels)
            + _cross_entropy_mean(logits_per_text, labels)) / 2


def gather_features_emulated(image_shards: Sequence[torch.Tensor],
                             text_shards: Sequence[torch.Tensor],
                             rank: int,
                             local_loss: bool,
                             gather_with_grad: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """What ``gather_features`` returns on ``rank`` (loss.py:46-63, non-horovod branch).

    gather_with_grad=True : every slot keeps its autograd link (loss.py:48-50).
    gather_with_grad=False: all slots are detached copies (loss.py:52-55); when
    ``local_loss`` is False the local slot is replaced by the grad-carrying local
    tensor (loss.py:56-59).  Row order is rank-major (torch.cat order, loss.py:60-61).
    """
    def build(shards):
        out = []
        for q, x in enumerate(shards):
            if gather_with_grad:
                out.append(x)
            elif q == rank and not local_loss:
                out.append(x)
            else:
                out.append(x.detach())
        return torch.cat(out, dim=0)
    return build(image_shards), build(text_shards)


def clip_loss_rank_block(image_shards: Sequence[torch.Tensor],
                         text_shards: Sequence[torch.Tensor],
                         logit_scale: torch.Tensor,
                         rank: int,
                         local_loss: bool,
                         gather_with_grad: bool) -> torch.Tensor:
    """Loss value computed by ``rank`` when world_size == len(shards) > 1
    (loss.py:103-113 for the logits, :93-94 for the label offset, :126-129 for CE)."""
    world = len(image_shards)
    assert world > 1
    all_i, all_t = gather_features_emulated(image_shards, text_shards, rank,
                                            local_loss, gather_with_grad)
    if local_loss:
        logits_per_image = (logit_scale * image_shards[rank]) @ all_t.T
        logits_per_text = (logit_scale * text_shards[rank]) @ all_i.T
    else:
        logits_per_image = (logit_scale * all_i) @ all_t.T
        logits_per_text = logits_per_image.T
    n = logits_per_image.shape[0]
    labels = torch.arange(n, dtype=torch.long)
    if local_loss:
        labels = labels + n * rank
    return (_cross_entropy_mean(logits_per_image, labels)
            + _cross_entropy_mean(logits_per_text, labels)) / 2


def clip_loss_all_ranks(image_shards: Sequence[torch.Tensor],
                        text_shards: Sequence[torch.Tensor],
                        logit_scale: float,
                        local_loss: bool,
                        gather_with_grad: bool,
                        dtype: torch.dtype = torch.float64):
    """Run every rank's forward + backward (emulated, see module docstring).

    Returns (losses[W], dI[W] list, dT[W] list, ds[W]) where dI[r]/dT[r] are the
    gradients rank r ends up with on its local features after the reference's
    backward, and ds[r] is rank r's logit_scale gradient (before DDP averaging).
    """
    world = len(image_shards)
    i_leaf = [x.detach().to(dtype).clone().requires_grad_(True) for x in image_shards]
    t_leaf = [x.detach().to(dtype).clone().requires_grad_(True) for x in text_shards]
    # one scale leaf per rank so that each rank's ds is separable
    s_leaf = [torch.tensor(float(logit_scale), dtype=dtype, requires_grad=True)
              for _ in range(world)]
    losses: List[torch.Tensor] = []
    if world == 1:
        losses.append(clip_loss_reference(i_leaf[0], t_leaf[0], s_leaf[0]))
    else:
        for r in range(world):
            losses.append(clip_loss_rank_block(i_leaf, t_leaf, s_leaf[r], r,
                                               local_loss, gather_with_grad))
    torch.stack(losses).sum().backward()
    zero = lambda x: torch.zeros_like(x) if x.grad is None else x.grad
    return ([l.detach() for l in losses],
            [zero(x) for x in i_leaf],
            [zero(x) for x in t_leaf],
            [zero(x) for x in s_leaf])


def clip_loss_fwd_bwd(image_features: torch.Tensor, text_features: torch.Tensor,
                      logit_scale: float, dtype: torch.dtype = torch.float32):
    """Single-rank convenience: (loss, dI, dT, ds) in ``dtype``."""
    losses, di, dt, ds = clip_loss_all_ranks([image_features], [text_features],
                                             logit_scale, False, False, dtype)
    return losses[0], di[0], dt[0], ds[0]


def clip_loss_row_block_sample(image_features: torch.Tensor,
                               text_features: torch.Tensor,
                               logit_scale: float,
                               rows: int) -> torch.Tensor:
    """Bounded CPU sample of the N x N workload used by bench.py's cpu_baseline:
    forward + backward of the loss terms owned by the first ``rows`` samples
    (row block of logits_per_image and of logits_per_text, i.e. what rank 0 of a
    local_loss run with n=rows would compute, loss.py:108-110).  Per-sample cost is the
    same as for the full batch, so samples/s = rows / time.  Returns the loss."""
    i_loc = image_features[:rows].detach().clone().requires_grad_(True)
    t_loc = text_features[:rows].detach().clone().requires_grad_(True)
    s = torch.tensor(float(logit_scale), dtype=image_features.dtype, requires_grad=True)
    logits_per_image = (s * i_loc) @ text_features.T
    logits_per_text = (s * t_loc) @ image_features.T
    labels = torch.arange(rows, dtype=torch.long)
    loss = (_cross_entropy_mean(logits_per_image, labels)
            + _cross_entropy_mean(logits_per_text, labels)) / 2
    loss.backward()
    return loss.detach()
