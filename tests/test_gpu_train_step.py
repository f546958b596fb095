"""latteclip_b200.train_step on the GPU: the DDP-safe step driven through a mock
model against the golden of the real ``train_one_epoch_v2`` loop, and the --accum-freq feature
cache against the golden of the reference ClipLoss in the accumulation pattern and the fp64 oracle."""

import math
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a = a.detach().double().cpu() if torch.is_tensor(a) else torch.as_tensor(np.asarray(a), dtype=torch.float64)
    b = b.detach().double().cpu() if torch.is_tensor(b) else torch.as_tensor(np.asarray(b), dtype=torch.float64)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


class TableModel(nn.Module):
    """Feature tables where the reference expects towers (as make_golden._TableModel)."""

    def __init__(self, img, cls_text, pimg, pgrp, bank, class_names, log_scale):
        super().__init__()
        self.img = nn.Parameter(img.clone())            # [nb, B, D]
        self.cls_text = nn.Parameter(cls_text.clone())
        self.pimg = nn.Parameter(pimg.clone())
        self.pgrp = nn.Parameter(pgrp.clone())
        self.logit_scale = nn.Parameter(torch.tensor(log_scale, dtype=img.dtype))
        self.memory_bank = nn.ParameterDict({c: nn.Parameter(bank[k].clone()) for k, c in enumerate(class_names)})
        self.class_names = class_names
        self.batch = 0

    def tokenizer(self, texts):
        tok = torch.zeros(len(texts), 2, dtype=torch.long)
        tok[:, 0] = torch.tensor([self.class_names.index(t.split("::")[1]) for t in texts])
        return tok

    def encode_image(self, images, normalize=True):
        return self.img[self.batch]

    def encode_text(self, tokens, normalize=True):
        kind, idx = int(tokens[0, 1]), tokens[:, 0]
        if kind == 0:
            return self.cls_text[idx]
        return (self.pimg if kind == 1 else self.pgrp)[self.batch][idx]


class Wrapper(nn.Module):
    """Stands in for DistributedDataParallel: attributes live on .module only."""

    def __init__(self, module):
        super().__init__()
        self.module = module


@pytest.mark.parametrize("name", ["b32_c7", "b64_c10_flags"])
def test_latteclip_step_matches_train_one_epoch_v2_golden(name):
    import latteclip_b200 as lb
    from latteclip_b200 import prototypes as P
    from latteclip_b200.train_step import ClassTextCache, latteclip_step
    g = load_golden(f"proto_step_{name}.npz")
    nb, flags = int(g["nb"]), [float(f) for f in g["flags"]]
    c, b = g["bank0"].shape[0], g["img"].shape[1]
    names = [f"class{j}" for j in range(c)]

    def t(x):
        return torch.from_numpy(np.asarray(x)).float()
    model = Wrapper(TableModel(t(g["img"]), t(g["cls_text"]), t(g["pimg"]), t(g["pgrp"]), t(g["bank0"]), names,
                               math.log(float(g["scale"])))).to(DEV)
    inner = model.module
    args = SimpleNamespace(device=DEV, alpha=float(g["alpha"]), use_image_caption=flags[0],
                           use_batch_caption=flags[1], use_template_caption=flags[2],
                           use_zeroshot_pseudolabel=flags[3], use_finetune_pseudolabel=flags[4], world_size=1)
    templates = [lambda cname: f"label::{cname}"]
    snapshot = P.stack_bank(inner.memory_bank, names)                 # train.py:347-350
    cache = ClassTextCache(inner.tokenizer, names, templates, DEV)
    loss_fn = lb.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True)
    for bi in range(nb):
        idx = torch.arange(b)
        pit = torch.zeros(b, 1, 2, dtype=torch.long); pit[:, 0, 0] = idx; pit[:, 0, 1] = 1    # noqa: E702
        pgt = torch.zeros(b, 1, 2, dtype=torch.long); pgt[:, 0, 0] = idx; pgt[:, 0, 1] = 2    # noqa: E702
        zcn = [(names[int(g["zs"][bi, i])],) for i in range(b)]
        batch = (torch.zeros(b, 1), torch.zeros(b, 1), torch.zeros(b, 1, 2), None, None, None, pit, pgt, None, zcn)
        for p in inner.parameters():
            p.grad = None
        inner.batch = bi
        out = latteclip_step(model, batch, loss_fn, args, names, templates, snapshot, class_text_cache=cache,
                             label_weight_axis="quirk")
        assert abs(float(out["contrastive_loss"]) - float(g[f"b{bi}_loss_ft"])) < 1e-4 * max(1.0, float(g[f"b{bi}_loss_ft"]))
        assert rel(inner.img.grad[bi], g[f"b{bi}_dI"]) < 1e-4
        assert rel(inner.cls_text.grad, g[f"b{bi}_dCls"]) < 1e-4
        assert rel(inner.pimg.grad[bi], g[f"b{bi}_dPimg"]) < 1e-4
        ref_dl = float(g[f"b{bi}_dlogscale"])
        assert abs(float(inner.logit_scale.grad) - ref_dl) < 1e-3 * max(1.0, abs(ref_dl))
    bank = torch.stack([inner.memory_bank[cn].detach() for cn in names])
    assert rel(bank, g["final_bank"]) < 1e-5
    assert all(isinstance(inner.memory_bank[cn], nn.Parameter) for cn in names)    # checkpoint format kept


@pytest.mark.parametrize("name", ["a3_m32", "a4_m40"])
@pytest.mark.parametrize("scale_grad", [True, False])
def test_feature_accumulator_fp32_matches_reference_golden(name, scale_grad):
    import latteclip_b200 as lb
    from latteclip_b200.train_step import FeatureAccumulator
    g = load_golden("clip_accum.npz")
    k, m, d, scale = g[f"{name}_meta"]
    k, m = int(k), int(m)
    i_all = torch.from_numpy(g[f"{name}_I"]).float().to(DEV)
    t_all = torch.from_numpy(g[f"{name}_T"]).float().to(DEV)
    acc = FeatureAccumulator(lb.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True), k)
    for j in range(k):
        acc.cache({"image_features": i_all[j * m:(j + 1) * m], "text_features": t_all[j * m:(j + 1) * m]})
    assert acc.ready() and acc._fused()
    for j in range(k):
        li = i_all[j * m:(j + 1) * m].clone().requires_grad_(True)
        lt = t_all[j * m:(j + 1) * m].clone().requires_grad_(True)
        s = torch.tensor(float(scale), device=DEV, requires_grad=scale_grad)
        losses = acc.micro_loss(j, {"image_features": li, "text_features": lt, "logit_scale": s})
        losses["loss"].backward()
        ref = float(g[f"{name}_loss{j}"])
        assert abs(float(losses["loss"]) - ref) < 1e-5 * abs(ref)         # north_star: fp32 loss <= 1e-5
        assert rel(li.grad, g[f"{name}_dI{j}"]) < 2e-5
        assert rel(lt.grad, g[f"{name}_dT{j}"]) < 2e-5
        if scale_grad:
            assert abs(float(s.grad) - float(g[f"{name}_ds{j}"])) < 2e-4 * abs(float(g[f"{name}_ds{j}"])) + 1e-7


@pytest.mark.parametrize("k,m,d,dtype", [(4, 256, 512, torch.bfloat16), (2, 1000, 768, torch.float16)])
@pytest.mark.parametrize("scale_grad", [True, False])
def test_feature_accumulator_16bit_matches_fp64_oracle(k, m, d, dtype, scale_grad):
    """Tensor-core path: the live micro-batch differs from its cached copy (as with dropout)."""
    import latteclip_b200 as lb
    from latteclip_b200.train_step import FeatureAccumulator
    from oracle.clip_loss import clip_loss_reference
    gen = torch.Generator().manual_seed(k * m)
    i_all = F.normalize(torch.randn(k * m, d, generator=gen), dim=1)
    t_all = F.normalize(i_all + 2.0 * torch.randn(k * m, d, generator=gen) / d ** 0.5, dim=1)
    i_dev, t_dev = i_all.to(DEV).to(dtype), t_all.to(DEV).to(dtype)
    acc = FeatureAccumulator(lb.ClipLoss(), k)
    for j in range(k):
        acc.cache({"image_features": i_dev[j * m:(j + 1) * m], "text_features": t_dev[j * m:(j + 1) * m]})
    scale = 60.0
    for j in (1, 0):
        sl = slice(j * m, (j + 1) * m)
        live_i = F.normalize(i_all[sl] + 0.05 * torch.randn(m, d, generator=gen), dim=1).to(DEV).to(dtype)
        live_t = F.normalize(t_all[sl] + 0.05 * torch.randn(m, d, generator=gen), dim=1).to(DEV).to(dtype)
        li, lt = live_i.clone().requires_grad_(True), live_t.clone().requires_grad_(True)
        s = torch.tensor(scale, device=DEV, requires_grad=scale_grad)
        losses = acc.micro_loss(j, {"image_features": li, "text_features": lt, "logit_scale": s})
        losses["loss"].backward()
        # fp64 on the same rounded values: cat(cached[:j] + [live] + cached[j+1:])  (train.py:1013-1015)
        ci = i_dev.detach().float().cpu().double().clone()
        ct = t_dev.detach().float().cpu().double().clone()
        ri = live_i.float().cpu().double().requires_grad_(True)
        rt = live_t.float().cpu().double().requires_grad_(True)
        sc = torch.tensor(scale, dtype=torch.float64, requires_grad=True)
        ref = clip_loss_reference(torch.cat([ci[:j * m], ri, ci[(j + 1) * m:]]),
                                  torch.cat([ct[:j * m], rt, ct[(j + 1) * m:]]), sc)
        ref.backward()
        assert abs(float(losses["loss"]) - float(ref)) < 2e-4 * abs(float(ref))
        gtol = 2.6e-3 if dtype == torch.bfloat16 else 2e-3       # bf16 output rounding: DESIGN section 4
        assert rel(li.grad, ri.grad) < gtol
        assert rel(lt.grad, rt.grad) < gtol
        if scale_grad:
            assert abs(float(s.grad) - float(sc.grad)) < 2e-3 * abs(float(sc.grad)) + 1e-7
