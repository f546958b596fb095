"""CPU tests of latteclip_b200.train_step: the --accum-freq feature cache and the
DDP-safe step.  The CUDA entry points are replaced by tests/_abi_double.py, so what is checked here
is the HOST logic (which blocks are rewritten, which backward runs, the gradient multiplier of the
block backward, the collectives of the 2-rank step); the kernels are covered by the -m gpu tests.
Expected values: tests/golden/clip_accum.npz, recorded from the reference's own ClipLoss driven in
the accumulation pattern of train.py:1002-1023 (single process and a real 2-rank gloo group)."""

import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

from conftest import load_golden  # noqa: E402


def _install_double():
    import _abi_double
    from latteclip_b200 import _lib
    _abi_double.install(_lib)
    from latteclip_b200 import train_step
    if torch.float64 not in train_step._FUSED_DTYPES:
        train_step._FUSED_DTYPES.append(torch.float64)      # the double computes in fp64


def rel(a, b):
    a = torch.as_tensor(np.asarray(a), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def test_oracle_accumulation_matches_reference_golden():
    from oracle.accum import clip_loss_accumulated
    g = load_golden("clip_accum.npz")
    for name in ("a3_m32", "a4_m40"):
        k, m, d, scale = g[f"{name}_meta"]
        k, m = int(k), int(m)
        i_all, t_all = torch.from_numpy(g[f"{name}_I"]), torch.from_numpy(g[f"{name}_T"])
        res = clip_loss_accumulated([i_all[j * m:(j + 1) * m] for j in range(k)],
                                    [t_all[j * m:(j + 1) * m] for j in range(k)], float(scale))
        for j, (loss, di, dt, ds) in enumerate(res):
            assert abs(float(loss) - float(g[f"{name}_loss{j}"])) < 1e-12 * abs(float(loss))
            assert rel(di, g[f"{name}_dI{j}"]) < 1e-12 and rel(dt, g[f"{name}_dT{j}"]) < 1e-12
            assert abs(float(ds) - float(g[f"{name}_ds{j}"])) < 1e-10 * abs(float(ds))


@pytest.mark.parametrize("name", ["a3_m32", "a4_m40"])
@pytest.mark.parametrize("scale_grad", [True, False])
def test_feature_accumulator_matches_reference_pattern(name, scale_grad):
    """scale_grad=True: the full backward runs and the live rows are sliced out; False: only the live
    row / column blocks are recomputed (grad_mult = m / N).  Both must give the reference's grads."""
    _install_double()
    import latteclip_b200 as lb
    from latteclip_b200.train_step import FeatureAccumulator
    g = load_golden("clip_accum.npz")
    k, m, d, scale = g[f"{name}_meta"]
    k, m = int(k), int(m)
    i_all, t_all = torch.from_numpy(g[f"{name}_I"]), torch.from_numpy(g[f"{name}_T"])
    acc = FeatureAccumulator(lb.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True), k)
    for j in range(k):
        acc.cache({"image_features": i_all[j * m:(j + 1) * m], "text_features": t_all[j * m:(j + 1) * m],
                   "logit_scale": torch.tensor(float(scale))})
    assert acc.ready() and acc._fused()
    order = list(range(k)) if name == "a3_m32" else [2, 0, 3, 1]      # any order restores the cache
    for j in order:
        # live features differ slightly from the cached ones in a real run (dropout); here they are
        # equal, as in the golden
        li = i_all[j * m:(j + 1) * m].clone().requires_grad_(True)
        lt = t_all[j * m:(j + 1) * m].clone().requires_grad_(True)
        s = torch.tensor(float(scale), dtype=torch.float64, requires_grad=scale_grad)
        losses = acc.micro_loss(j, {"image_features": li, "text_features": lt, "logit_scale": s})
        assert set(losses) == {"contrastive_loss", "loss"}
        losses["loss"].backward()
        assert abs(float(losses["loss"]) - float(g[f"{name}_loss{j}"])) < 1e-6 * float(g[f"{name}_loss{j}"])
        assert rel(li.grad, g[f"{name}_dI{j}"]) < 1e-6
        assert rel(lt.grad, g[f"{name}_dT{j}"]) < 1e-6
        if scale_grad:
            assert abs(float(s.grad) - float(g[f"{name}_ds{j}"])) < 1e-5 * abs(float(g[f"{name}_ds{j}"]))
        else:
            assert s.grad is None


def test_feature_accumulator_detects_stale_backward():
    _install_double()
    import latteclip_b200 as lb
    from latteclip_b200.train_step import FeatureAccumulator
    torch.manual_seed(0)
    acc = FeatureAccumulator(lb.ClipLoss(), 2)
    feats = [F.normalize(torch.randn(8, 16), dim=1) for _ in range(4)]
    for j in range(2):
        acc.cache({"image_features": feats[2 * j], "text_features": feats[2 * j + 1]})
    outs = []
    for j in range(2):
        li, lt = feats[2 * j].clone().requires_grad_(True), feats[2 * j + 1].clone().requires_grad_(True)
        outs.append(acc.micro_loss(j, {"image_features": li, "text_features": lt,
                                       "logit_scale": torch.tensor(10.0)})["loss"])
    with pytest.raises(RuntimeError, match="modified by an inplace operation"):
        outs[0].backward()          # micro-batch 1 has overwritten the work buffers
    outs[1].backward()


# ------------------------------------------------------------------------------ 2-rank gloo
def _accum_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _install_double()
    import latteclip_b200 as lb
    from latteclip_b200 import _lib
    from latteclip_b200.train_step import FeatureAccumulator
    _lib.rank_sweep_supported = lambda dtype, dim: False
    g = np.load(os.path.join(HERE, "golden", "clip_accum.npz"))
    k, m, d, scale, _ = g["w2_meta"]
    k, m = int(k), int(m)
    i_all, t_all = torch.from_numpy(g["w2_I"]), torch.from_numpy(g["w2_T"])
    base = rank * k * m
    loss_fn = lb.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True, rank=rank,
                          world_size=world)
    # forward() wants CUDA tensors only through _lib; the double accepts CPU ones
    acc = FeatureAccumulator(loss_fn, k)
    for j in range(k):
        acc.cache({"image_features": i_all[base + j * m: base + (j + 1) * m],
                   "text_features": t_all[base + j * m: base + (j + 1) * m]})
    assert not acc._fused()       # multi-rank: the reference's concatenation, then the gathered loss
    out = {}
    for j in range(k):
        li = i_all[base + j * m: base + (j + 1) * m].clone().requires_grad_(True)
        lt = t_all[base + j * m: base + (j + 1) * m].clone().requires_grad_(True)
        s = torch.tensor(float(scale), dtype=torch.float64, requires_grad=True)
        losses = acc.micro_loss(j, {"image_features": li, "text_features": lt, "logit_scale": s})
        losses["loss"].backward()
        out[j] = dict(loss=float(losses["loss"]), dI=li.grad.numpy(), dT=lt.grad.numpy(), ds=float(s.grad))
    ret[rank] = out
    dist.barrier()
    dist.destroy_process_group()


def test_feature_accumulator_two_ranks_matches_gloo_reference():
    ret = mp.Manager().dict()
    mp.spawn(_accum_worker, args=(2, 29761, ret), nprocs=2, join=True)
    g = load_golden("clip_accum.npz")
    k = int(g["w2_meta"][0])
    for r in range(2):
        for j in range(k):
            o = ret[r][j]
            assert abs(o["loss"] - float(g[f"w2_r{r}_loss{j}"])) < 1e-6 * abs(float(g[f"w2_r{r}_loss{j}"]))
            # the ABI (and its double) hands the LSE vectors between the calls in fp32
            assert rel(o["dI"], g[f"w2_r{r}_dI{j}"]) < 2e-5
            assert rel(o["dT"], g[f"w2_r{r}_dT{j}"]) < 2e-5
            assert abs(o["ds"] - float(g[f"w2_r{r}_ds{j}"])) < 1e-4 * abs(float(g[f"w2_r{r}_ds{j}"]))


# ------------------------------------------------------------------------------ DDP-safe step
class TableModel(nn.Module):
    """Feature tables where the reference expects towers (same idea as make_golden._TableModel)."""

    def __init__(self, img, cls_text, pimg, pgrp, bank, class_names, log_scale):
        super().__init__()
        self.img = nn.Parameter(img.clone())
        self.cls_text = nn.Parameter(cls_text.clone())
        self.pimg = nn.Parameter(pimg.clone())
        self.pgrp = nn.Parameter(pgrp.clone())
        self.logit_scale = nn.Parameter(torch.tensor(log_scale, dtype=img.dtype))
        self.memory_bank = nn.ParameterDict({c: nn.Parameter(bank[k].clone()) for k, c in enumerate(class_names)})
        self.class_names = class_names

    def tokenizer(self, texts):
        tok = torch.zeros(len(texts), 2, dtype=torch.long)
        tok[:, 0] = torch.tensor([self.class_names.index(t.split("::")[1]) for t in texts])
        return tok

    def encode_image(self, images, normalize=True):
        return self.img[images[:, 0].long()]

    def encode_text(self, tokens, normalize=True):
        kind, idx = int(tokens[0, 1]), tokens[:, 0]
        return (self.cls_text, self.pimg, self.pgrp)[kind][idx]


class Wrapper(nn.Module):
    """Stands in for DistributedDataParallel: the attributes live on .module only."""

    def __init__(self, module):
        super().__init__()
        self.module = module


def make_step_problem(b_total, d, c, seed):
    g = torch.Generator().manual_seed(seed)
    names = [f"class{j}" for j in range(c)]
    bank0 = F.normalize(torch.randn(c, d, generator=g, dtype=torch.float64), dim=1)
    cls_text = F.normalize(bank0 + 0.3 * torch.randn(c, d, generator=g, dtype=torch.float64), dim=1)
    true_cls = torch.randint(0, c, (b_total,), generator=g)
    mk = lambda s: F.normalize(bank0[true_cls] + s * torch.randn(b_total, d, generator=g, dtype=torch.float64), dim=1)  # noqa: E731
    img, pimg, pgrp = mk(0.4), mk(0.3), mk(0.25)
    zs = torch.where(torch.rand(b_total, generator=g) < 0.7, true_cls, torch.randint(0, c, (b_total,), generator=g))
    return names, bank0, cls_text, img, pimg, pgrp, zs


def make_batch(rows, zs, names):
    b = len(rows)
    images = rows.double()[:, None]
    pit = torch.zeros(b, 1, 2, dtype=torch.long); pit[:, 0, 0] = rows; pit[:, 0, 1] = 1    # noqa: E702
    pgt = torch.zeros(b, 1, 2, dtype=torch.long); pgt[:, 0, 0] = rows; pgt[:, 0, 1] = 2    # noqa: E702
    zcn = [(names[int(zs[i])],) for i in rows]
    return (images, torch.zeros(b, 1), torch.zeros(b, 1, 2), None, None, None, pit, pgt, None, zcn)


def step_args(world, rank):
    return SimpleNamespace(device="cpu", alpha=0.05, use_image_caption=1.0, use_batch_caption=1.0,
                           use_template_caption=1.0, use_zeroshot_pseudolabel=0.5,
                           use_finetune_pseudolabel=1.0, world_size=world, rank=rank)


def run_step(world, rank, rows, problem):
    import latteclip_b200 as lb
    from latteclip_b200 import prototypes as P
    from latteclip_b200.train_step import latteclip_step
    names, bank0, cls_text, img, pimg, pgrp, zs = problem
    model = Wrapper(TableModel(img, cls_text, pimg, pgrp, bank0, names, float(np.log(30.0))))
    assert not hasattr(model, "memory_bank")        # what breaks the reference under DDP
    snapshot = P.stack_bank(model.module.memory_bank, names).double()
    loss = lb.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True, rank=rank, world_size=world)
    orig = lb.ClipLoss.forward

    def forward(self, image_features, text_features, logit_scale, output_dict=False):
        total = lb.loss._FusedClipLoss.apply(image_features, text_features, logit_scale, self.local_loss,
                                             self.gather_with_grad, self.rank, self.world_size, None,
                                             image_features.dtype, False)
        return {"contrastive_loss": total} if output_dict else total
    lb.ClipLoss.forward = forward                   # skip the CUDA-only autocast query
    try:
        out = latteclip_step(model, make_batch(rows, zs, names), loss, step_args(world, rank), names,
                             [lambda cname: f"label::{cname}"], snapshot, label_weight_axis="row")
    finally:
        lb.ClipLoss.forward = orig
    m = model.module
    grads = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None and "memory_bank" not in n}
    bank = torch.stack([m.memory_bank[c].detach() for c in names])
    return float(out["loss"]), grads, bank


def _step_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _install_double()
    from latteclip_b200 import _lib
    _lib.rank_sweep_supported = lambda dtype, dim: False
    problem = make_step_problem(32, 24, 6, 11)
    rows = torch.arange(rank * 16, (rank + 1) * 16)
    loss, grads, bank = run_step(world, rank, rows, problem)
    ret[rank] = dict(loss=loss, grads={k: v.numpy() for k, v in grads.items()}, bank=bank.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_safe_step_two_ranks_equals_single_process_on_the_concatenated_batch():
    """Two ranks, each with half of the batch behind a DDP-like wrapper: the averaged tower gradients,
    the mean loss and the memory bank must equal ONE process stepping on the concatenated batch
    (ClipLoss(local_loss, gather_with_grad) differentiates W x the global-mean loss per rank, DDP
    divides by W; update_bank all-reduces the class sums)."""
    ret = mp.Manager().dict()
    mp.spawn(_step_worker, args=(2, 29771, ret), nprocs=2, join=True)
    _install_double()
    problem = make_step_problem(32, 24, 6, 11)
    loss1, grads1, bank1 = run_step(1, 0, torch.arange(32), problem)
    assert abs((ret[0]["loss"] + ret[1]["loss"]) / 2 - loss1) < 1e-6 * abs(loss1)
    for r in range(2):
        assert rel(ret[r]["bank"], bank1) < 1e-6                      # identical banks on every rank
        for name, gref in grads1.items():
            assert rel(ret[r]["grads"][name], gref) < 1e-5, name      # rank-averaged == single process
    assert np.array_equal(ret[0]["bank"], ret[1]["bank"])
