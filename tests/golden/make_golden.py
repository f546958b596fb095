#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container (needs /root/reference mounted):

    python tests/golden/make_golden.py

What is executed is the reference's own code, imported from /root/reference/src
(only ``ftfy`` is stubbed, it is not used on this path):

  * ``open_clip.loss.ClipLoss`` forward + autograd backward, world_size 1
      -> clip_w1_*.npz
  * ``open_clip.loss.ClipLoss`` with world_size 2 and 4 on real gloo process groups,
    all four (local_loss, gather_with_grad) combinations
      -> clip_dist_w{2,4}.npz
  * ``training.train.compute_text_weights``
      -> text_margins.npz
  * one/two full iterations of ``training.train.train_one_epoch_v2`` (the inline
    prototype / pseudo-label / mixture / EMA / bank-update code, train.py:384-530)
    driven through mock model / data / optimizer objects that only supply feature
    tables, so every arithmetic statement on the hot path is the reference's
      -> proto_step_*.npz
  * ``training.zero_shot.accuracy`` / ``run`` and ``training.train.accuracy``
      -> zero_shot_eval.npz
  * ``open_clip.loss.SigLipLoss`` forward + backward, world size 1 and with its ring exchange
    on real gloo process groups of 2, 3 and 4 ranks
      -> siglip.npz
  * the ``--accum-freq`` feature-cache pattern of upstream open_clip (present in the reference at
    train.py:972-1024 behind ``raise NotImplemented()``) driven with the reference's ``ClipLoss``
      -> clip_accum.npz
  * ``open_clip.loss.DistillClipLoss`` forward + backward (loss.py:324-362), world size 1
      -> distill.npz

The fixtures hold both the inputs and the reference outputs, so tests never need the
reference at run time (it does not exist on the GPU box).
"""

import math
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"


def load_reference():
    if not os.path.isdir(REF_SRC):
        raise SystemExit("reference not mounted at /root/reference")
    sys.modules.setdefault("ftfy", types.ModuleType("ftfy"))
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import open_clip  # noqa
    import training.train as tt  # noqa
    return open_clip, tt


def synth_pairs(n, d, sigma, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    i = F.normalize(torch.randn(n, d, generator=g), dim=1)
    # sigma is in units of the feature norm: sigma=1 -> cos(I_i, T_i) ~ 0.7
    t = F.normalize(i + sigma * torch.randn(n, d, generator=g) / math.sqrt(d), dim=1)
    return i.to(dtype), t.to(dtype)


# ----------------------------------------------------------------------------------
# 1. ClipLoss, world_size == 1
# ----------------------------------------------------------------------------------
def gen_clip_w1(open_clip):
    cases = [
        # name, N, D, sigma, scale, text_norm_jitter
        ("small_s100", 96, 64, 3.0, 100.0, 0.0),
        ("small_s14", 96, 64, 2.0, 1.0 / 0.07, 0.0),
        ("ragged_s100", 77, 48, 1.5, 100.0, 0.01),   # N, D not multiples of a tile
        ("cfg1_s100", 256, 512, 4.0, 100.0, 0.005),  # BASELINE config 1 shape, non-unit text
    ]
    for k, (name, n, d, sigma, scale, jitter) in enumerate(cases):
        i, t = synth_pairs(n, d, sigma, 1234 + k)
        if jitter:
            g = torch.Generator().manual_seed(99 + k)
            t = t * (1.0 - jitter * torch.rand(n, 1, generator=g))   # |T| ~ 0.995
        out = {}
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            il = i.to(dt).clone().requires_grad_(True)
            tl = t.to(dt).clone().requires_grad_(True)
            log_s = torch.tensor(math.log(scale), dtype=dt, requires_grad=True)
            s = log_s.exp()
            s.retain_grad()
            loss_mod = open_clip.ClipLoss(cache_labels=True)
            res = loss_mod(il, tl, s, output_dict=True)
            assert list(res.keys()) == ["contrastive_loss"]
            res["contrastive_loss"].backward()
            out[f"loss_{tag}"] = res["contrastive_loss"].detach().numpy()
            # grads are stored in float32 to keep the fixtures small (loss/ds stay f64)
            out[f"dI_{tag}"] = il.grad.numpy().astype(np.float32)
            out[f"dT_{tag}"] = tl.grad.numpy().astype(np.float32)
            out[f"ds_{tag}"] = s.grad.numpy()
        np.savez_compressed(os.path.join(HERE, f"clip_w1_{name}.npz"),
                            I=i.numpy(), T=t.numpy(), scale=np.float64(scale), **out)
        print("clip_w1", name, float(out["loss_f64"]))


# ----------------------------------------------------------------------------------
# 2. ClipLoss on real gloo process groups
# ----------------------------------------------------------------------------------
def _dist_worker(rank, world, port, i_all, t_all, scale, ret):
    import torch.distributed as dist
    sys.modules.setdefault("ftfy", types.ModuleType("ftfy"))
    sys.path.insert(0, REF_SRC)
    import open_clip
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = i_all.shape[0] // world
    res = {}
    for local_loss in (False, True):
        for gwg in (False, True):
            il = i_all[rank * n:(rank + 1) * n].clone().requires_grad_(True)
            tl = t_all[rank * n:(rank + 1) * n].clone().requires_grad_(True)
            s = torch.tensor(scale, dtype=i_all.dtype, requires_grad=True)
            mod = open_clip.ClipLoss(local_loss=local_loss, gather_with_grad=gwg,
                                     cache_labels=True, rank=rank, world_size=world)
            loss = mod(il, tl, s)
            loss.backward()
            key = f"ll{int(local_loss)}_gwg{int(gwg)}"
            res[key] = dict(loss=loss.detach().numpy(), dI=il.grad.numpy(),
                            dT=tl.grad.numpy(), ds=s.grad.numpy())
    ret[rank] = res
    dist.barrier()
    dist.destroy_process_group()


def gen_clip_dist():
    import torch.multiprocessing as mp
    for world, port in ((2, 29611), (4, 29613)):
        n_glob, d, scale = 64, 32, 100.0
        i, t = synth_pairs(n_glob, d, 1.5, 4321 + world, torch.float64)
        mgr = mp.Manager()
        ret = mgr.dict()
        mp.spawn(_dist_worker, args=(world, port, i, t, scale, ret), nprocs=world, join=True)
        out = {"I": i.numpy(), "T": t.numpy(), "scale": np.float64(scale), "world": np.int64(world)}
        for r in range(world):
            for key, v in ret[r].items():
                for nm, arr in v.items():
                    out[f"{key}_r{r}_{nm}"] = arr
        np.savez_compressed(os.path.join(HERE, f"clip_dist_w{world}.npz"), **out)
        print("clip_dist", world, float(ret[0]["ll1_gwg1"]["loss"]))


# ----------------------------------------------------------------------------------
# 3. compute_text_weights
# ----------------------------------------------------------------------------------
def gen_text_margins(tt):
    g = torch.Generator().manual_seed(777)
    protos = F.normalize(torch.randn(47, 512, generator=g), dim=1) * 0.997
    x = F.normalize(protos[torch.randint(0, 47, (300,), generator=g)]
                    + 0.6 * torch.randn(300, 512, generator=g) / math.sqrt(512) * 4, dim=1)
    preds = torch.randint(0, 47, (300,), generator=g)
    m32 = tt.compute_text_weights(x, protos, preds)
    m64 = tt.compute_text_weights(x.double(), protos.double(), preds)
    np.savez_compressed(os.path.join(HERE, "text_margins.npz"), X=x.numpy(), P=protos.numpy(),
                        margin_f32=m32.numpy(), margin_f64=m64.numpy())
    print("text_margins", float(m64.mean()))


# ----------------------------------------------------------------------------------
# 4. train_one_epoch_v2 through mocks
# ----------------------------------------------------------------------------------
class _TableModel(nn.Module):
    """Supplies feature tables where the reference expects towers.  Everything the
    reference does to those features is the reference's own code."""

    def __init__(self, img, cls_text, pimg, pgrp, bank, class_names, log_scale):
        super().__init__()
        self.img = nn.Parameter(img.clone())            # [nb, B, D]
        self.cls_text = nn.Parameter(cls_text.clone())  # [C, D]
        self.pimg = nn.Parameter(pimg.clone())          # [nb, B, D]
        self.pgrp = nn.Parameter(pgrp.clone())          # [nb, B, D]
        self.logit_scale = nn.Parameter(torch.tensor(log_scale, dtype=img.dtype))
        self.memory_bank = nn.ParameterDict(
            {c: nn.Parameter(bank[k].clone()) for k, c in enumerate(class_names)})
        self.class_names = class_names
        self.batch = 0

    def tokenizer(self, texts):
        ids = [self.class_names.index(t.split("::")[1]) for t in texts]
        tok = torch.zeros(len(ids), 2, dtype=torch.long)
        tok[:, 0] = torch.tensor(ids)
        tok[:, 1] = 0
        return tok

    def encode_image(self, images, normalize=True):
        return self.img[self.batch]

    def encode_text(self, tokens, normalize=True):
        kind = int(tokens[0, 1])
        idx = tokens[:, 0]
        if kind == 0:
            return self.cls_text[idx]
        if kind == 1:
            return self.pimg[self.batch][idx]
        return self.pgrp[self.batch][idx]


class _SnapOpt:
    """Optimizer stand-in: records gradients at step(), applies no update."""

    def __init__(self, model):
        self.model = model
        self.param_groups = [{"lr": 0.0}]
        self.snaps = []

    def zero_grad(self):
        for p in self.model.parameters():
            p.grad = None

    def step(self):
        m = self.model
        b = m.batch
        self.snaps.append(dict(
            dI=m.img.grad[b].clone(), dCls=m.cls_text.grad.clone(),
            dPimg=m.pimg.grad[b].clone(), dPgrp=m.pgrp.grad[b].clone(),
            dlogscale=m.logit_scale.grad.clone()))
        m.batch += 1


class _SpyLoss(nn.Module):
    def __init__(self, inner):
        super().__init__()
        self.inner = inner
        self.calls = []

    def forward(self, image_features, text_features, logit_scale, output_dict=False):
        out = self.inner(image_features, text_features, logit_scale, output_dict=output_dict)
        self.calls.append(dict(text=text_features.detach().clone(),
                               loss=out["contrastive_loss"].detach().clone()))
        return out


def gen_proto_step(open_clip, tt):
    cases = [
        # name, B(=D), C, batches, scale, alpha, flags(image, batch, template, zs, ft), dtype
        ("b32_c7", 32, 7, 2, 100.0, 0.01, (1.0, 1.0, 1.0, 1.0, 1.0), torch.float64),
        ("b64_c10", 64, 10, 1, 100.0, 0.01, (1.0, 1.0, 1.0, 1.0, 1.0), torch.float32),
        ("b64_c10_flags", 64, 10, 1, 1.0 / 0.07, 0.05, (1.0, 0.0, 1.0, 0.5, 1.0), torch.float64),
    ]
    for k, (name, b, c, nb, scale, alpha, flags, dt) in enumerate(cases):
        d = b  # the reference's broadcast at train.py:476 only runs when B == D
        g = torch.Generator().manual_seed(2024 + k)
        class_names = [f"class{j}" for j in range(c)]
        bank0 = F.normalize(torch.randn(c, d, generator=g), dim=1)
        cls_text = F.normalize(bank0 + 0.3 * torch.randn(c, d, generator=g) / math.sqrt(d) * 3, dim=1)
        true_cls = torch.randint(0, c, (nb, b), generator=g)
        img = F.normalize(bank0[true_cls] + 1.2 * torch.randn(nb, b, d, generator=g) / math.sqrt(d) * 3, dim=2)
        pimg = F.normalize(bank0[true_cls] + 0.9 * torch.randn(nb, b, d, generator=g) / math.sqrt(d) * 3, dim=2)
        pgrp = F.normalize(bank0[true_cls] + 0.7 * torch.randn(nb, b, d, generator=g) / math.sqrt(d) * 3, dim=2)
        zs = torch.where(torch.rand(nb, b, generator=g) < 0.7, true_cls,
                         torch.randint(0, c, (nb, b), generator=g))
        bank0, cls_text, img, pimg, pgrp = (x.to(dt) for x in (bank0, cls_text, img, pimg, pgrp))

        model = _TableModel(img, cls_text, pimg, pgrp, bank0, class_names, math.log(scale))
        opt = _SnapOpt(model)
        spy = _SpyLoss(open_clip.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True))

        def make_batch(bi):
            idx = torch.arange(b)
            texts = torch.zeros(b, 1, 2, dtype=torch.long)
            pit = torch.zeros(b, 1, 2, dtype=torch.long); pit[:, 0, 0] = idx; pit[:, 0, 1] = 1
            pgt = torch.zeros(b, 1, 2, dtype=torch.long); pgt[:, 0, 0] = idx; pgt[:, 0, 1] = 2
            zcn = [(class_names[int(zs[bi, i])],) for i in range(b)]
            return (torch.zeros(b, 1), torch.zeros(b, 1), texts, None, None, None, pit, pgt, None, zcn)

        class _DL(list):
            num_batches = nb
            num_samples = nb * b
        dl = _DL([make_batch(bi) for bi in range(nb)])
        key = "synth-train-zero-shot-classification"
        data = {"train": SimpleNamespace(set_epoch=lambda e: None, dataloader=dl),
                key: SimpleNamespace(class_names=class_names,
                                     templates=[lambda cname: f"label::{cname}"])}
        args = SimpleNamespace(
            device="cpu", precision="fp32", zeroshot_eval_data="synth",
            extract_features_split="train", distill=False, accum_freq=1, skip_scheduler=True,
            lr_scheduler="cosine", alpha=alpha, use_image_caption=flags[0],
            use_batch_caption=flags[1], use_template_caption=flags[2],
            use_zeroshot_pseudolabel=flags[3], use_finetune_pseudolabel=flags[4],
            horovod=False, grad_clip_norm=None, log_every_n_steps=1000, rank=0, local_rank=0,
            world_size=1, batch_size=b, wandb=False)

        banks = [bank0.clone()]
        # run batch by batch so the bank after every batch can be recorded
        tt.train_one_epoch_v2(model, data, spy, 0, opt, None, None, None, args, None)
        # (the reference loop consumed all nb batches; per-batch bank states are
        #  recovered below for nb == 1, and for nb == 2 the final bank is recorded)
        final_bank = torch.stack([model.memory_bank[cn].detach() for cn in class_names])

        out = dict(bank0=bank0.numpy(), cls_text=cls_text.numpy(), img=img.numpy(),
                   pimg=pimg.numpy(), pgrp=pgrp.numpy(), zs=zs.numpy(),
                   scale=np.float64(scale), alpha=np.float64(alpha),
                   flags=np.array(flags, dtype=np.float64), nb=np.int64(nb),
                   final_bank=final_bank.numpy())
        for bi in range(nb):
            out[f"b{bi}_t_ft"] = spy.calls[2 * bi]["text"].numpy()
            out[f"b{bi}_t_zs"] = spy.calls[2 * bi + 1]["text"].numpy()
            out[f"b{bi}_loss_ft"] = spy.calls[2 * bi]["loss"].numpy()
            out[f"b{bi}_loss_zs"] = spy.calls[2 * bi + 1]["loss"].numpy()
            for nm, v in opt.snaps[bi].items():
                out[f"b{bi}_{nm}"] = v.numpy()
        np.savez_compressed(os.path.join(HERE, f"proto_step_{name}.npz"), **out)
        print("proto_step", name, [float(cl["loss"]) for cl in spy.calls])


def gen_zero_shot(tt):
    """training.zero_shot.accuracy / run and training.train.accuracy (zero_shot.py:14-52,
    train.py:1128-1138) on seeded features; the 'model' only hands the features through."""
    import training.zero_shot as zs
    g = torch.Generator().manual_seed(4242)
    b, d, c = 192, 64, 23
    protos = F.normalize(torch.randn(c, d, generator=g), dim=1)
    target = torch.randint(0, c, (b,), generator=g)
    feats = F.normalize(protos[target] + 1.2 * torch.randn(b, d, generator=g) / math.sqrt(d) * 3.0, dim=1)
    classifier = protos.T.contiguous()
    logits = 100.0 * feats @ classifier
    accs_zs = zs.accuracy(logits, target, topk=(1, 5, 10))
    accs_tt, top_logits, top_ids = tt.accuracy(logits, target, topk=(1, 5, 10))

    class _Tower(nn.Module):
        def forward(self, image=None):
            return {"image_features": image}

    batches = [(None, feats[k:k + 50], target[k:k + 50]) for k in range(0, b, 50)]   # ragged tail
    args = SimpleNamespace(precision="fp32", device="cpu", batch_size=50)
    rates = zs.run(_Tower(), classifier, batches, args)
    np.savez_compressed(os.path.join(HERE, "zero_shot_eval.npz"), feats=feats.numpy(),
                        classifier=classifier.numpy(), target=target.numpy(),
                        accs_zero_shot=np.asarray(accs_zs), accs_train=np.asarray(accs_tt),
                        top_logits=top_logits.numpy(), top_ids=top_ids.numpy(),
                        rates=np.asarray(rates), batch=np.asarray(50))
    print("zero_shot_eval", accs_zs, accs_tt, rates)


# ----------------------------------------------------------------------------------
# 6. SigLipLoss (loss.py:453-560), world size 1 and on real gloo rings (world 2, 3, 4)
# ----------------------------------------------------------------------------------
def _siglip_once(open_clip, il, tl, scale, bias, rank, world):
    il = il.clone().requires_grad_(True)
    tl = tl.clone().requires_grad_(True)
    s = torch.tensor(scale, dtype=il.dtype, requires_grad=True)
    b = torch.tensor(bias, dtype=il.dtype, requires_grad=True)
    mod = open_clip.loss.SigLipLoss(rank=rank, world_size=world)
    loss = mod(il, tl, s, b)
    loss.backward()
    return dict(loss=loss.detach().numpy(), dI=il.grad.numpy(), dT=tl.grad.numpy(),
                ds=s.grad.numpy(), db=b.grad.numpy())


def _siglip_worker(rank, world, port, i_all, t_all, scale, bias, ret):
    import torch.distributed as dist
    sys.modules.setdefault("ftfy", types.ModuleType("ftfy"))
    sys.path.insert(0, REF_SRC)
    import open_clip
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = i_all.shape[0] // world
    ret[rank] = _siglip_once(open_clip, i_all[rank * n:(rank + 1) * n], t_all[rank * n:(rank + 1) * n],
                             scale, bias, rank, world)
    dist.barrier()
    dist.destroy_process_group()


# ----------------------------------------------------------------------------------
# 7. --accum-freq feature cache (upstream code kept in the reference at train.py:972-1024)
# ----------------------------------------------------------------------------------
def gen_clip_accum(open_clip):
    """The accumulation pattern of train.py:1002-1023 driven with the reference's own ClipLoss:
    per micro-batch j the loss over cat(cached blocks, live block j), one backward each."""
    out = {}
    for name, k, m, d, sigma, scale in (("a3_m32", 3, 32, 64, 3.0, 100.0), ("a4_m40", 4, 40, 48, 5.0, 1.0 / 0.07)):
        i_all, t_all = synth_pairs(k * m, d, sigma, 900 + k + m, torch.float64)
        loss_fn = open_clip.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True)
        cached_i = [i_all[j * m:(j + 1) * m] for j in range(k)]
        cached_t = [t_all[j * m:(j + 1) * m] for j in range(k)]
        out[f"{name}_I"], out[f"{name}_T"] = i_all.numpy(), t_all.numpy()
        out[f"{name}_meta"] = np.array([k, m, d, scale], dtype=np.float64)
        for j in range(k):
            live_i = cached_i[j].clone().requires_grad_(True)
            live_t = cached_t[j].clone().requires_grad_(True)
            s = torch.tensor(scale, dtype=torch.float64, requires_grad=True)
            inputs = {"image_features": torch.cat(cached_i[:j] + [live_i] + cached_i[j + 1:]),
                      "text_features": torch.cat(cached_t[:j] + [live_t] + cached_t[j + 1:])}
            losses = loss_fn(**inputs, logit_scale=s, output_dict=True)
            total = sum(losses.values())
            total.backward()
            out[f"{name}_loss{j}"] = total.detach().numpy()
            out[f"{name}_dI{j}"] = live_i.grad.numpy()
            out[f"{name}_dT{j}"] = live_t.grad.numpy()
            out[f"{name}_ds{j}"] = s.grad.numpy()
        print("clip_accum", name, float(total))
    # the same pattern on a real 2-rank gloo group: every rank caches k micro-batches of m rows and
    # the reference ClipLoss gathers the concatenated [k*m, D] features of both ranks
    import torch.multiprocessing as mp
    world, k, m, d, scale = 2, 2, 16, 32, 100.0
    i_all, t_all = synth_pairs(world * k * m, d, 2.0, 977, torch.float64)
    ret = mp.Manager().dict()
    mp.spawn(_accum_dist_worker, args=(world, 29641, i_all, t_all, k, m, scale, ret), nprocs=world, join=True)
    out["w2_I"], out["w2_T"] = i_all.numpy(), t_all.numpy()
    out["w2_meta"] = np.array([k, m, d, scale, world], dtype=np.float64)
    for r in range(world):
        for key, arr in ret[r].items():
            out[f"w2_r{r}_{key}"] = arr
    print("clip_accum w2", float(ret[0]["loss0"]))
    np.savez_compressed(os.path.join(HERE, "clip_accum.npz"), **out)


def _accum_dist_worker(rank, world, port, i_all, t_all, k, m, scale, ret):
    """Rank r holds rows [r*k*m, (r+1)*k*m) as k micro-batches; train.py:1002-1023 per micro-batch."""
    import torch.distributed as dist
    sys.modules.setdefault("ftfy", types.ModuleType("ftfy"))
    sys.path.insert(0, REF_SRC)
    import open_clip
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    loss_fn = open_clip.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True,
                                 rank=rank, world_size=world)
    base = rank * k * m
    cached_i = [i_all[base + j * m: base + (j + 1) * m] for j in range(k)]
    cached_t = [t_all[base + j * m: base + (j + 1) * m] for j in range(k)]
    res = {}
    for j in range(k):
        live_i = cached_i[j].clone().requires_grad_(True)
        live_t = cached_t[j].clone().requires_grad_(True)
        s = torch.tensor(scale, dtype=torch.float64, requires_grad=True)
        losses = loss_fn(image_features=torch.cat(cached_i[:j] + [live_i] + cached_i[j + 1:]),
                         text_features=torch.cat(cached_t[:j] + [live_t] + cached_t[j + 1:]),
                         logit_scale=s, output_dict=True)
        total = sum(losses.values())
        total.backward()
        res[f"loss{j}"] = total.detach().numpy()
        res[f"dI{j}"], res[f"dT{j}"], res[f"ds{j}"] = live_i.grad.numpy(), live_t.grad.numpy(), s.grad.numpy()
    ret[rank] = res
    dist.barrier()
    dist.destroy_process_group()


# ----------------------------------------------------------------------------------
# 8. DistillClipLoss (loss.py:324-362)
# ----------------------------------------------------------------------------------
def gen_distill(open_clip):
    from open_clip.loss import DistillClipLoss
    out = {}
    for name, n, d, sig_s, sig_t, s_s, s_t in (("n96_d64", 96, 64, 3.0, 1.5, 30.0, 100.0),
                                                ("n200_d128", 200, 128, 5.0, 2.0, 1.0 / 0.07, 50.0),
                                                ("n64_d32_close", 64, 32, 2.0, 2.0, 40.0, 40.0)):
        i_s, t_s = synth_pairs(n, d, sig_s, 5000 + n, torch.float64)
        if name.endswith("close"):        # a student close to its teacher: the gradient is a small difference
            g = torch.Generator().manual_seed(77)
            i_t = F.normalize(i_s + 0.02 * torch.randn(n, d, generator=g, dtype=torch.float64), dim=1)
            t_t = F.normalize(t_s + 0.02 * torch.randn(n, d, generator=g, dtype=torch.float64), dim=1)
        else:
            i_t, t_t = synth_pairs(n, d, sig_t, 6000 + n, torch.float64)
        out[f"{name}_I"], out[f"{name}_T"] = i_s.numpy(), t_s.numpy()
        out[f"{name}_It"], out[f"{name}_Tt"] = i_t.numpy(), t_t.numpy()
        out[f"{name}_scales"] = np.array([s_s, s_t], dtype=np.float64)
        for tag, dt in (("f64", torch.float64), ("f32", torch.float32)):
            il = i_s.to(dt).clone().requires_grad_(True)
            tl = t_s.to(dt).clone().requires_grad_(True)
            s = torch.tensor(s_s, dtype=dt, requires_grad=True)
            mod = DistillClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True)
            res = mod(il, tl, s, i_t.to(dt), t_t.to(dt), torch.tensor(s_t, dtype=dt), output_dict=True)
            assert res["contrastive_loss"] == 0
            res["distill_loss"].backward()
            out[f"{name}_loss_{tag}"] = res["distill_loss"].detach().numpy()
            out[f"{name}_dI_{tag}"], out[f"{name}_dT_{tag}"] = il.grad.numpy(), tl.grad.numpy()
            out[f"{name}_ds_{tag}"] = s.grad.numpy()
        print("distill", name, float(out[f"{name}_loss_f64"]))
    np.savez_compressed(os.path.join(HERE, "distill.npz"), **out)


def gen_siglip(open_clip):
    import torch.multiprocessing as mp
    out = {}
    # world size 1: the init operating point (s = 10, b = -10, main.py:225-227) and a trained-like one
    for name, n, d, sigma, scale, bias in (("init", 96, 64, 1.5, 10.0, -10.0),
                                           ("hot", 130, 40, 3.0, 80.0, -12.0)):
        i, t = synth_pairs(n, d, sigma, 900 + n, torch.float64)
        i, t = i.bfloat16().double(), t.bfloat16().double()      # exactly representable in bf16
        r = _siglip_once(open_clip, i, t, scale, bias, 0, 1)
        r32 = _siglip_once(open_clip, i.float(), t.float(), scale, bias, 0, 1)
        out.update({f"{name}_I": i.numpy(), f"{name}_T": t.numpy(), f"{name}_scale": np.float64(scale),
                    f"{name}_bias": np.float64(bias)})
        out.update({f"{name}_{k}_f64": v for k, v in r.items()})
        out.update({f"{name}_{k}_f32": v for k, v in r32.items()})
        print("siglip", name, float(r["loss"]), float(r32["loss"]))
    for world, port in ((2, 29641), (3, 29643), (4, 29645)):
        n_glob, d, scale, bias = 24 * world, 32, 20.0, -6.0
        i, t = synth_pairs(n_glob, d, 2.0, 977 + world, torch.float64)
        i, t = i.bfloat16().double(), t.bfloat16().double()
        mgr = mp.Manager()
        ret = mgr.dict()
        mp.spawn(_siglip_worker, args=(world, port, i, t, scale, bias, ret), nprocs=world, join=True)
        out.update({f"w{world}_I": i.numpy(), f"w{world}_T": t.numpy(), f"w{world}_scale": np.float64(scale),
                    f"w{world}_bias": np.float64(bias)})
        for r in range(world):
            for k, v in ret[r].items():
                out[f"w{world}_r{r}_{k}"] = v
        print("siglip ring", world, [float(ret[r]["loss"]) for r in range(world)])
    np.savez_compressed(os.path.join(HERE, "siglip.npz"), **out)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    open_clip, tt = load_reference()
    gen_clip_w1(open_clip)
    gen_text_margins(tt)
    gen_proto_step(open_clip, tt)
    gen_zero_shot(tt)
    gen_siglip(open_clip)
    gen_clip_dist()
    gen_clip_accum(open_clip)
    gen_distill(open_clip)


if __name__ == "__main__":
    main()
