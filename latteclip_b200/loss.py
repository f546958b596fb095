"""Drop-in for the ClipLoss part of open_clip's ``loss.py``.

Mirrors the reference interface (same names, arguments, defaults and error behaviour):
  * ``gather_features``  -- /root/reference/src/open_clip/loss.py:19-63
  * ``ClipLoss``         -- loss.py:66-130 (``__init__`` :68-87, ``get_ground_truth`` :89-100,
                            ``get_logits`` :102-118, ``forward`` :120-130)
  * ``create_loss``      -- /root/reference/src/open_clip/factory.py:323-351 (ClipLoss and SigLipLoss
                            branches; ``SigLipLoss`` itself is in ``siglip.py``)

``ClipLoss.forward`` is the hot path: it never materialises the logit matrices.  It calls
the CUDA extension (C ABI in include/latte_b200.h) through ``latteclip_b200._lib`` inside a
``torch.autograd.Function``; the cross-rank exchanges are listed in ``_FusedClipLoss``.  ``get_logits`` /
``get_ground_truth`` remain as materialising utilities for subclasses
(DistillClipLoss, loss.py:341-345); they are plain torch and not on the fused path.
"""

from __future__ import annotations

import os

import torch
import torch.nn as nn

try:
    import torch.distributed as dist
    has_distributed = True
except ImportError:  # pragma: no cover
    dist = None
    has_distributed = False

from . import _lib


# ------------------------------------------------------------------------------------------
# differentiable all-gather used by the materialising utility path (loss.py:49-50)
# ------------------------------------------------------------------------------------------
class _AllGatherWithGrad(torch.autograd.Function):
    """all_gather whose backward is reduce-scatter(SUM), like torch.distributed.nn.all_gather."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        world = dist.get_world_size(group)
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)
        return out

    @staticmethod
    def backward(ctx, grad):
        world = dist.get_world_size(ctx.group)
        rank = dist.get_rank(ctx.group)
        grad = grad.contiguous()
        n = grad.shape[0] // world
        if dist.get_backend(ctx.group) == "nccl":
            out = torch.empty((n,) + tuple(grad.shape[1:]), dtype=grad.dtype, device=grad.device)
            dist.reduce_scatter_tensor(out, grad, op=dist.ReduceOp.SUM, group=ctx.group)
        else:  # gloo has no reduce_scatter
            dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=ctx.group)
            out = grad[rank * n:(rank + 1) * n].clone()
        return out, None


def _all_gather_cat(x: torch.Tensor, group=None) -> torch.Tensor:
    world = dist.get_world_size(group)
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def gather_features(
        image_features,
        text_features,
        local_loss=False,
        gather_with_grad=False,
        rank=0,
        world_size=1,
        use_horovod=False
):
    """Same contract as the reference (loss.py:19-63): returns (all_image_features,
    all_text_features), rank-major row order.  gather_with_grad=True keeps the autograd link
    (backward = reduce-scatter); False gathers without grad and, unless local_loss, re-inserts
    the grad-carrying local shard (loss.py:56-59)."""
    assert has_distributed, 'torch.distributed did not import correctly, please use a PyTorch version with support.'
    if use_horovod:
        # the reference's horovod branch (loss.py:29-45) needs horovod, which this build does not ship
        raise NotImplementedError("latteclip_b200: horovod gather is not supported; use torch.distributed")
    if gather_with_grad:
        all_image_features = _AllGatherWithGrad.apply(image_features, None)
        all_text_features = _AllGatherWithGrad.apply(text_features, None)
    else:
        with torch.no_grad():
            all_image_features = _all_gather_cat(image_features)
            all_text_features = _all_gather_cat(text_features)
        if not local_loss:
            n = image_features.shape[0]
            chunks_i = list(all_image_features.split(n, dim=0))
            chunks_t = list(all_text_features.split(n, dim=0))
            chunks_i[rank] = image_features
            chunks_t[rank] = text_features
            all_image_features = torch.cat(chunks_i, dim=0)
            all_text_features = torch.cat(chunks_t, dim=0)
    return all_image_features, all_text_features


# ------------------------------------------------------------------------------------------
# fused ClipLoss
# ------------------------------------------------------------------------------------------
def _reduce_scatter_sum(x: torch.Tensor, n: int, rank: int, group=None) -> torch.Tensor:
    """rows [rank*n, (rank+1)*n) of the sum over ranks of x [W*n, D] (the backward of the
    reference's differentiable all-gather, torch/distributed/nn/functional.py:343-354)."""
    if dist.get_backend(group) == "nccl":
        out = torch.empty((n,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.reduce_scatter_tensor(out, x.contiguous(), op=dist.ReduceOp.SUM, group=group)
        return out
    x = x.contiguous()      # gloo has no reduce_scatter
    dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
    return x[rank * n:(rank + 1) * n].clone()


# fp32 text-gradient accumulators mapped into every rank of the group (torch symmetric memory):
# the gradient GEMM of each rank adds its rows straight into their owner's accumulator over
# NVLink, which fuses the reduce-scatter of the reference's all_gather backward into the GEMM.
_PEER_ACC = {}


def _bwd_sweeps(world_size: int) -> int:
    """Recompute sweeps per rank in the multi-rank backward: 1 (default) = rows only, the text-side
    product is reduce-scattered inside the GEMM; 2 = rows and columns, nothing is exchanged
    (LATTE_B200_BWD_SWEEPS=2; executes 12 n N D for the 6 credited)."""
    env = os.environ.get("LATTE_B200_BWD_SWEEPS", "")
    if env in ("1", "2"):
        return int(env)
    return 1


def _peer_accumulator(n: int, dim: int, device, group):
    """-> (acc [n, dim] fp32, symmetric-memory handle) or None when peer mapping is unavailable
    (non-NCCL backend, no P2P, LATTE_B200_NO_P2P=1); rendezvous happens once per shape."""
    if os.environ.get("LATTE_B200_NO_P2P") == "1" or device.type != "cuda":
        return None
    key = (n, dim, device.index, id(group))
    if key in _PEER_ACC:
        return _PEER_ACC[key]
    entry = None
    try:
        if dist.get_backend(group) == "nccl" and dist.get_world_size(group) <= 8:
            import torch.distributed._symmetric_memory as symm
            acc = symm.empty(n, dim, dtype=torch.float32, device=device)
            hdl = symm.rendezvous(acc, group if group is not None else dist.group.WORLD)
            entry = (acc, hdl, [int(p) for p in hdl.buffer_ptrs])
    except Exception:      # no symmetric memory on this system: NCCL reduce-scatter path
        entry = None
    # the decision must be the same on every rank (the collectives differ)
    ok = torch.tensor([1 if entry is not None else 0], device=device, dtype=torch.int32)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok) == 0:
        entry = None
    _PEER_ACC[key] = entry
    return entry


# ------------------------------------------------------------------------------------------
# Peer-memory exchange state (torch symmetric memory: one allocation per rank, mapped by all)
# ------------------------------------------------------------------------------------------
# Per (shard shape, dtype, group) every rank owns `_COMM_SLOTS` slots, each holding the gathered
# feature buffer [2, N, D], the forward payload blocks, the fp32 text-gradient accumulator and a
# flag block (include/latte_b200.h: latte_comm_t).  The kernels store into / add into the peers'
# slots and synchronise with generation-numbered flags; no host-side barrier or collective runs in
# a step.  A slot is busy from a forward to the end of its backward (the gathered text features are
# re-read by the recompute sweep); a forward that finds no free slot falls back to NCCL -- every
# rank takes the same decision because the call sequence is the same.
_COMM_STATES = {}
_COMM_SLOTS = 4


def _round_up(x: int, a: int) -> int:
    return (x + a - 1) // a * a


class _CommSlot:
    def __init__(self, state, index: int, offset: int):
        self.state, self.index, self.offset = state, index, offset
        self.gen = 0
        self.busy = False
        st = state
        base = st.buf[offset:offset + st.slot_bytes]
        self.gather = base[:st.gather_bytes].view(st.dtype).view(2, st.n_all, st.dim)
        # per-rank pointers of the four regions
        self._ptrs = [[p + offset + off for p in st.base_ptrs]
                      for off in (0, st.off_payload, st.off_acc, st.off_flags)]
        self.comm = _lib.make_comm(st.rank, st.world, 1, self._ptrs[0], self._ptrs[1], self._ptrs[2],
                                   self._ptrs[3], st.payload_stride)

    @property
    def all_img(self):
        return self.gather[0]

    @property
    def all_txt(self):
        return self.gather[1]

    def acquire(self):
        self.gen += 1
        self.comm.gen = self.gen
        self.busy = True
        return self

    def release(self, signal: bool):
        """Mark the slot free; ``signal`` launches the release kernel (a forward without backward:
        the backward's last kernel normally publishes the release itself)."""
        if not self.busy:
            return
        if signal:
            _lib.comm_release(self.comm, self.state.device)
        self.busy = False


class _SlotLease:
    """Ties a slot to the autograd node that saved its gathered features: if the node dies without a
    backward (a logged validation loss, an exception), the slot is released when the node is
    collected instead of staying busy forever."""

    def __init__(self, slot):
        self.slot = slot
        self.gen = slot.gen

    def done(self):
        self.slot = None

    def __del__(self):
        slot = self.slot
        if slot is not None and slot.busy and slot.gen == self.gen:
            try:
                slot.release(signal=True)
            except Exception:      # interpreter shutdown
                pass


class _CommState:
    def __init__(self, n, dim, dtype, device, group, world, rank, buf, hdl):
        self.n, self.dim, self.dtype, self.device = n, dim, dtype, device
        self.world, self.rank = world, rank
        self.n_all = n * world
        self.buf, self.hdl = buf, hdl
        self.base_ptrs = [int(p) for p in hdl.buffer_ptrs]
        (self.gather_bytes, self.off_payload, self.off_acc, self.off_flags, self.slot_bytes,
         self.payload_stride) = _CommState.layout(n, dim, world)
        self.slots = [_CommSlot(self, k, k * self.slot_bytes) for k in range(_COMM_SLOTS)]
        self.next = 0

    @staticmethod
    def layout(n, dim, world):
        n_all = n * world
        gather_bytes = 2 * n_all * dim * 2
        stride = _round_up(2 * n_all + 3 * n, 4)
        payload_bytes = (world * stride + world * 2 * n_all) * 4
        off_payload = _round_up(gather_bytes, 256)
        off_acc = off_payload + _round_up(payload_bytes, 256)
        off_flags = off_acc + _round_up(n * dim * 4, 256)
        slot_bytes = off_flags + _round_up(_lib.COMM_FLAG_INTS * 4, 256)
        return gather_bytes, off_payload, off_acc, off_flags, slot_bytes, stride

    def acquire(self):
        for k in range(len(self.slots)):
            sl = self.slots[(self.next + k) % len(self.slots)]
            if not sl.busy:
                self.next = (sl.index + 1) % len(self.slots)
                return sl.acquire()
        return None


def _comm_state(n: int, dim: int, dtype, device, group, world: int, rank: int):
    """-> _CommState or None (no symmetric memory / LATTE_B200_NO_P2P=1 / unsupported shape); the
    rendezvous happens once per shape and the decision is agreed on by all ranks."""
    if os.environ.get("LATTE_B200_NO_P2P") == "1" or device.type != "cuda":
        return None
    if dtype not in (torch.bfloat16, torch.float16) or (n * dim * 2) % 16 != 0 or dim % 4 != 0:
        return None
    if world > _lib.COMM_MAX_RANKS:
        return None
    key = (n, dim, dtype, device.index, id(group))
    if key in _COMM_STATES:
        return _COMM_STATES[key]
    state = None
    try:
        if dist.get_backend(group) == "nccl":
            import torch.distributed._symmetric_memory as symm
            total = _CommState.layout(n, dim, world)[4] * _COMM_SLOTS
            buf = symm.empty(total, dtype=torch.uint8, device=device)
            buf.zero_()                                   # flags and accumulators start at zero
            hdl = symm.rendezvous(buf, group if group is not None else dist.group.WORLD)
            state = _CommState(n, dim, dtype, device, group, world, rank, buf, hdl)
    except (ImportError, RuntimeError, AttributeError) as exc:
        if os.environ.get("LATTE_B200_VERBOSE"):
            print(f"latteclip_b200: peer-memory exchange unavailable ({exc}); using NCCL", flush=True)
        state = None
    # the decision must be the same on every rank (the collectives differ); this all-reduce also
    # orders every rank's zero-fill before anybody's first signal
    ok = torch.tensor([1 if state is not None else 0], device=device, dtype=torch.int32)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok) == 0:
        state = None
    _COMM_STATES[key] = state
    return state


class _FusedClipLoss(torch.autograd.Function):
    """loss.py:102-130 fused.  Gradient contract:
         local_loss & gather_with_grad : grads = d(sum_r L_r)/dx_local   (W x global-mean grad)
         !local_loss & gather_with_grad: loss = L_global on every rank, same feature grads
         !local_loss & !gather_with_grad: grads = 1 x local slice of dL_global
         local_loss & !gather_with_grad : own-block terms only (gathered copies carry no grad)

    Multi-rank, 16-bit features with dim <= 768 (every mode but the last): ONE logit sweep per
    rank in each direction, exchanges by the kernels themselves over NVLink peer memory:
      forward : a store kernel writes this rank's text shard into every rank's gathered buffer
                (the all-gather of loss.py:49-50; the images of other ranks are never needed);
                the sweep of own images x all texts starts on the shards as they land and yields
                the row LSEs and per-column (max, sum) partials; each rank stores that payload
                ([2N + 3n] floats) into every peer's block and merges the blocks into every
                column's LSE as soon as they arrived.
      backward: one recompute sweep -> G[rows of this rank, :]; d_img is local; the text gradient
                G^T.img_loc covers ALL columns and must be reduce-scattered -- the reference's own
                collective (the backward of its all_gather): the GEMM epilogue adds every row
                straight into its owner rank's fp32 accumulator (system-scope red over NVLink), so
                the exchange overlaps the GEMM tile by tile; a last kernel waits for every rank's
                "adds done" flag and casts the accumulator.
    d loss / d logit_scale then covers this rank's rows x all columns: a different partition over
    ranks of the same global sum as the reference's (identical after DDP's all-reduce of the
    parameter gradient -- logit_scale.grad must be all-reduced to match the reference).
    Without symmetric memory (or LATTE_B200_NO_P2P=1) the same flow runs over NCCL: all-gathers of
    the features and of the payload, fp32 [N, D] partial + reduce_scatter.
    Other cases (fp32 features, local_loss without gather_with_grad): each rank sweeps its row block
    and its column block and the backward exchanges only the LSE vectors.
    """

    @staticmethod
    def forward(ctx, image_features, text_features, logit_scale, local_loss, gather_with_grad,
                rank, world_size, group, compute_dtype, normalize):
        # Operands: on the tensor-core path ONE fused kernel per matrix normalises (opt-in), rounds
        # to the compute dtype and stores fp16 -- no ATen cast, no per-step copies in the backward;
        # fp32 compute keeps the features as they are (SIMT parity kernels).
        raw_i, raw_t = image_features.detach(), text_features.detach()
        inv_i = inv_t = None
        if compute_dtype in (torch.bfloat16, torch.float16) and \
                _lib.rank_sweep_supported(compute_dtype, raw_i.shape[1]):
            img, inv_i = _lib.prep_features(raw_i, compute_dtype, normalize)
            txt, inv_t = _lib.prep_features(raw_t, compute_dtype, normalize)
        else:
            if normalize:
                raise NotImplementedError(
                    "latteclip_b200.ClipLoss(normalize_features=True) needs the tensor-core path "
                    "(16-bit compute dtype, dim <= 768, dim % 8 == 0)")
            img, txt = raw_i.to(compute_dtype), raw_t.to(compute_dtype)
        ctx.norm = (raw_i, raw_t, inv_i, inv_t) if normalize else None
        cross_terms = not (world_size > 1 and local_loss and not gather_with_grad)
        rank_sweep = False
        slot = None
        lease = None
        two_sweeps = False
        if world_size > 1:
            label_offset = rank * img.shape[0]
            rank_sweep = cross_terms and _lib.rank_sweep_supported(img.dtype, img.shape[1])
            two_sweeps = rank_sweep and _bwd_sweeps(world_size) == 2
            if rank_sweep and img.dtype == txt.dtype:
                state = _comm_state(img.shape[0], img.shape[1], img.dtype, img.device, group,
                                    world_size, rank)
                slot = state.acquire() if state is not None else None
            if slot is not None:
                # one NVLink store kernel instead of NCCL all-gathers; the images of the other ranks
                # are only needed when the backward recomputes rows AND columns
                _lib.comm_push(slot.comm, txt, img if two_sweeps else None,
                               tensor_stride_bytes=slot.all_img.numel() * slot.all_img.element_size())
                all_img, all_txt = (slot.all_img if two_sweeps else None), slot.all_txt
            else:
                all_img = _all_gather_cat(img, group) if (not rank_sweep or two_sweeps) else None
                all_txt = _all_gather_cat(txt, group)
        else:
            all_img, all_txt, label_offset = img, txt, 0
        if rank_sweep and slot is not None:
            row_lse_all, row_nll_all, col_lse_all, col_nll_all, loss, lse_stats, _ = _lib.clip_fwd_rank(
                slot.comm, img, all_txt, label_offset, logit_scale)
            stats = (row_lse_all, col_lse_all, row_nll_all, col_nll_all, lse_stats)
        elif rank_sweep:
            payload = _lib.clip_fwd_rows(img, all_txt, label_offset, logit_scale)
            gathered = _all_gather_cat(payload.reshape(1, -1), group)            # [W, 2N + 3n]
            row_lse_all, row_nll_all, col_lse_all, col_nll_all, loss, lse_stats = _lib.clip_fwd_cols(
                gathered, all_img if all_img is not None else _all_gather_cat(img, group), all_txt,
                img.shape[0], label_offset, logit_scale)
            stats = (row_lse_all, col_lse_all, row_nll_all, col_nll_all, lse_stats)
        else:
            row_lse, col_lse, loss, row_nll, col_nll, lse_stats = _lib.clip_fwd(
                img, txt, all_img, all_txt, label_offset, logit_scale, with_nll=True, with_stats=True)
            stats = (row_lse, col_lse, row_nll, col_nll, lse_stats)
        loss = loss.reshape(())
        if world_size > 1 and not local_loss:
            # L_global = mean over ranks of the per-rank block losses (equal shard sizes)
            loss = loss.clone()
            dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
            loss = loss / world_size
        if slot is not None:
            if any(ctx.needs_input_grad[:3]):
                lease = _SlotLease(slot)
            else:
                slot.release(signal=True)                 # no backward will come for this call
                slot = None
        ctx.lease = lease
        if all_img is None:
            all_img = img                                 # placeholder: never read in the one-sweep modes
        ctx.save_for_backward(img, txt, all_img, all_txt, logit_scale.detach(), *stats)
        ctx.cfg = (local_loss, gather_with_grad, rank, world_size, group, label_offset, rank_sweep,
                   two_sweeps)
        ctx.scale_meta = (logit_scale.dtype, logit_scale.shape)
        ctx.feat_dtypes = (image_features.dtype, text_features.dtype)
        return loss

    @staticmethod
    def _grad_dtype(ctx):
        if ctx.norm is not None:
            return torch.float32          # the normalisation's backward rounds once, at its end
        di, dt = ctx.feat_dtypes
        return di if di == dt and di in (torch.float32, torch.bfloat16, torch.float16) else torch.float32

    @staticmethod
    def backward(ctx, grad_out):
        img, txt, all_img, all_txt, scale, row_lse, col_lse, row_nll, col_nll, lse_stats = ctx.saved_tensors
        (local_loss, gather_with_grad, rank, world_size, group, label_offset, rank_sweep,
         two_sweeps) = ctx.cfg
        if world_size > 1 and not rank_sweep:
            lse_stats = None    # per-rank statistics: the kernels recompute them from the gathered vectors
        cross_terms = not (world_size > 1 and local_loss and not gather_with_grad)
        grad_mult = 1.0
        if world_size > 1 and not local_loss and not gather_with_grad:
            grad_mult = 1.0 / world_size
        gdt = _FusedClipLoss._grad_dtype(ctx)
        lease = getattr(ctx, "lease", None)
        slot = lease.slot if lease is not None else None
        if lease is not None and (slot is None or not slot.busy or slot.gen != lease.gen):
            raise RuntimeError(
                "latteclip_b200.ClipLoss: backward called twice on a multi-rank forward (the gathered "
                "features of that step were released after the first backward); call forward again")
        if rank_sweep and two_sweeps:
            # rows AND columns recomputed per rank, nothing exchanged (LATTE_B200_BWD_SWEEPS=2)
            d_img, d_txt, d_scale = _lib.clip_bwd(
                img, txt, all_img, all_txt, label_offset, scale, row_lse, col_lse, grad_out,
                grad_mult, True, grad_dtype=gdt, row_nll_all=row_nll, col_nll_all=col_nll, lse_stats=lse_stats)
            if slot is not None:
                slot.release(signal=True)
        elif rank_sweep and slot is not None:
            # fused reduce-scatter: every rank's GEMM adds into the owners' accumulators; the call's
            # last kernel publishes the slot's release
            d_img, d_txt, d_scale = _lib.clip_bwd(
                img, txt, None, all_txt, label_offset, scale, row_lse, col_lse, grad_out,
                grad_mult, True, grad_dtype=gdt, row_nll_all=row_nll, col_nll_all=col_nll,
                comm=slot.comm, lse_stats=lse_stats)
            slot.release(signal=False)
        elif rank_sweep:
            d_img, d_part, d_scale = _lib.clip_bwd(
                img, txt, None, all_txt, label_offset, scale, row_lse, col_lse, grad_out,
                grad_mult, True, grad_dtype=gdt, row_nll_all=row_nll, col_nll_all=col_nll, partial=True,
                lse_stats=lse_stats)
            d_txt = _reduce_scatter_sum(d_part, img.shape[0], rank, group).to(gdt)
        else:
            if world_size > 1:
                four = _all_gather_cat(torch.stack([row_lse, col_lse, row_nll, col_nll], dim=1), group)
                row_all, col_all, rown_all, coln_all = (four[:, k].contiguous() for k in range(4))
            else:
                row_all, col_all, rown_all, coln_all = row_lse, col_lse, row_nll, col_nll
            d_img, d_txt, d_scale = _lib.clip_bwd(
                img, txt, all_img, all_txt, label_offset, scale, row_all, col_all, grad_out,
                grad_mult, cross_terms, grad_dtype=gdt, row_nll_all=rown_all, col_nll_all=coln_all,
                lse_stats=lse_stats)
        if lease is not None:
            lease.done()
        if world_size > 1 and not local_loss:
            # every rank differentiates the same L_global: d/ds is the rank mean of the block sums
            d_scale = d_scale / grad_mult
            dist.all_reduce(d_scale, op=dist.ReduceOp.SUM, group=group)
            d_scale = d_scale / world_size
        if ctx.norm is not None:
            raw_i, raw_t, inv_i, inv_t = ctx.norm
            d_img = _lib.normalize_bwd(d_img, raw_i, inv_i)
            d_txt = _lib.normalize_bwd(d_txt, raw_t, inv_t)
        di_t, dt_t = ctx.feat_dtypes
        if d_img.dtype != di_t:
            d_img = d_img.to(di_t)
        if d_txt.dtype != dt_t:
            d_txt = d_txt.to(dt_t)
        s_dtype, s_shape = ctx.scale_meta
        d_scale = d_scale.reshape(s_shape).to(s_dtype)
        need = ctx.needs_input_grad
        return (d_img if need[0] else None, d_txt if need[1] else None,
                d_scale if need[2] else None, None, None, None, None, None, None, None)


class ClipLoss(nn.Module):
    """Same constructor, methods and call as the reference's ``ClipLoss`` (loss.py:66-130).

    Notes beyond the reference: (1) scratch -- the backward keeps the fp16 gradient weights
    ``G [n, N]`` (n * N * 2 bytes per rank: 2 GiB at N = 32768 on one GPU) in ONE growable buffer per
    device and stream (``latteclip_b200.clear_workspace_cache()`` releases it); the logits themselves are
    never stored.  (2) feature gradients are not bit-reproducible from run to run (fp32 ``red.add`` of
    the stream-K partial tiles in arrival order: last-bit differences); the loss is.  (3) multi-rank
    ``local_loss``: ``logit_scale.grad`` is this rank's partition (own rows x all columns) of the global
    sum -- all-reduce it, as DDP does for the parameter, to match the reference.  (4)
    ``normalize_features=True`` (extension, default off) fuses ``F.normalize`` of both inputs."""

    def __init__(
            self,
            local_loss=False,
            gather_with_grad=False,
            cache_labels=False,
            rank=0,
            world_size=1,
            use_horovod=False,
            normalize_features=False,
    ):
        super().__init__()
        # Extension (not in the reference signature; default off, SURVEY fact 4: ClipLoss does not
        # normalise): L2-normalise both feature matrices inside the fused operand-preparation
        # kernel, i.e. what encode_image / encode_text(normalize=True) do in the towers
        # (model.py:415-418, 420-437), with the matching backward.
        self.normalize_features = normalize_features
        self.local_loss = local_loss
        self.gather_with_grad = gather_with_grad
        self.cache_labels = cache_labels
        self.rank = rank
        self.world_size = world_size
        self.use_horovod = use_horovod
        self.group = None          # process group for the collectives (None = default group)

        # cache state (kept for interface parity, loss.py:85-87)
        self.prev_num_logits = 0
        self.labels = {}

    def get_ground_truth(self, device, num_logits) -> torch.Tensor:
        # loss.py:89-100 -- on the fused path the label is implicit (column rank*n + i)
        if self.prev_num_logits != num_logits or device not in self.labels:
            labels = torch.arange(num_logits, device=device, dtype=torch.long)
            if self.world_size > 1 and self.local_loss:
                labels = labels + num_logits * self.rank
            if self.cache_labels:
                self.labels[device] = labels
                self.prev_num_logits = num_logits
        else:
            labels = self.labels[device]
        return labels

    def get_logits(self, image_features, text_features, logit_scale):
        # loss.py:102-118 -- materialising utility, NOT used by forward()
        if self.world_size > 1:
            all_image_features, all_text_features = gather_features(
                image_features, text_features,
                self.local_loss, self.gather_with_grad, self.rank, self.world_size, self.use_horovod)
            if self.local_loss:
                logits_per_image = logit_scale * image_features @ all_text_features.T
                logits_per_text = logit_scale * text_features @ all_image_features.T
            else:
                logits_per_image = logit_scale * all_image_features @ all_text_features.T
                logits_per_text = logits_per_image.T
        else:
            logits_per_image = logit_scale * image_features @ text_features.T
            logits_per_text = logit_scale * text_features @ image_features.T
        return logits_per_image, logits_per_text

    def forward(self, image_features, text_features, logit_scale, output_dict=False):
        if self.use_horovod:
            raise NotImplementedError("latteclip_b200: horovod is not supported")
        if image_features.shape != text_features.shape:
            raise RuntimeError(
                f"image_features {tuple(image_features.shape)} and text_features "
                f"{tuple(text_features.shape)} must have the same shape")
        if not torch.is_tensor(logit_scale):
            logit_scale = torch.tensor(float(logit_scale), device=image_features.device)
        # One compute dtype for both operands: the autocast dtype when autocast is on (the
        # reference's matmuls run in it), otherwise the promoted dtype of the two inputs.
        if torch.is_autocast_enabled():
            cdt = torch.get_autocast_dtype("cuda")
        else:
            cdt = torch.promote_types(image_features.dtype, text_features.dtype)
        if cdt not in (torch.float32, torch.bfloat16, torch.float16):
            cdt = torch.float32
        with torch.autocast(device_type="cuda", enabled=False):
            total_loss = _FusedClipLoss.apply(
                image_features, text_features, logit_scale,
                self.local_loss, self.gather_with_grad, self.rank, self.world_size, self.group,
                cdt, self.normalize_features)
        return {"contrastive_loss": total_loss} if output_dict else total_loss


def create_loss(args):
    """The reference factory (factory.py:323-351): SigLipLoss on ``args.siglip`` (:337-342), else
    ClipLoss (:344-351).  CoCaLoss is outside the accelerated path."""
    if "coca" in getattr(args, "model", "").lower():
        raise NotImplementedError(
            "latteclip_b200.create_loss provides ClipLoss and SigLipLoss; use the reference factory "
            "for CoCaLoss (it wraps a ClipLoss, loss.py:309, which can be this one)")
    # (the reference factory has no distillation branch: args.distill still gets a ClipLoss here;
    #  latteclip_b200.DistillClipLoss is constructed directly, like open_clip.loss.DistillClipLoss)
    if getattr(args, "siglip", False):
        assert not args.horovod, "Horovod not currently supported for SigLip"
        from .siglip import SigLipLoss
        return SigLipLoss(rank=args.rank, world_size=args.world_size)
    return ClipLoss(
        local_loss=args.local_loss,
        gather_with_grad=args.gather_with_grad,
        cache_labels=True,
        rank=args.rank,
        world_size=args.world_size,
        use_horovod=args.horovod,
    )
