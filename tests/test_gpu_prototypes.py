"""GPU parity tests for the prototype / pseudo-label / mixture / EMA / bank kernels, against
the CPU oracle and the golden fixtures recorded from the unmodified train_one_epoch_v2."""

import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _f64(x):
    if torch.is_tensor(x):
        return x.detach().to(torch.float64).cpu()
    return torch.as_tensor(np.asarray(x), dtype=torch.float64)


def rel(a, b):
    a, b = _f64(a), _f64(b)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def proto_inputs(b, d, c, seed):
    g = torch.Generator().manual_seed(seed)
    bank = F.normalize(torch.randn(c, d, generator=g), dim=1)
    cls_text = F.normalize(bank + 0.9 * torch.randn(c, d, generator=g) / math.sqrt(d), dim=1)
    true_cls = torch.randint(0, c, (b,), generator=g)
    img = F.normalize(bank[true_cls] + 3.0 * torch.randn(b, d, generator=g) / math.sqrt(d), dim=1)
    pimg = F.normalize(bank[true_cls] + 2.0 * torch.randn(b, d, generator=g) / math.sqrt(d), dim=1)
    pgrp = F.normalize(bank[true_cls] + 1.5 * torch.randn(b, d, generator=g) / math.sqrt(d), dim=1)
    zs = torch.where(torch.rand(b, generator=g) < 0.7, true_cls, torch.randint(0, c, (b,), generator=g))
    return bank, cls_text, img, pimg, pgrp, zs


# ------------------------------------------------------------------ K7: classifier
@pytest.mark.parametrize("c,d", [(47, 512), (397, 768), (10, 33)])
def test_build_classifier(c, d):
    from latteclip_b200 import prototypes as P
    import oracle
    x = torch.randn(c, d) * 3
    got = P.build_classifier(x.to(DEV))
    assert rel(got, oracle.build_classifier(x)) < 1e-6


# ------------------------------------------------------------------ K8: pseudo-labels (bit exact)
@pytest.mark.parametrize("b,c,d", [(1000, 100, 512), (10000, 397, 512), (5000, 1000, 768), (100000, 100, 512)])
def test_pseudo_labels_bit_exact_outside_ties(b, c, d):
    """BASELINE config 5.  A row counts as a tie when its fp64 top-1/top-2 gap is <= tau
    (tau = 1e-6 * scale); everywhere else the labels must be identical."""
    from latteclip_b200 import prototypes as P
    bank, _, img, _, _, _ = proto_inputs(b, d, c, 31 + c)
    clf = F.normalize(bank, dim=1)
    got = P.pseudo_label(img.to(DEV), clf.to(DEV), 100.0).cpu()
    logits = 100.0 * img.double() @ clf.double().T
    top2 = logits.topk(2, dim=1)
    ref = top2.indices[:, 0]
    gap = top2.values[:, 0] - top2.values[:, 1]
    clear = gap > 1e-6 * 100.0
    assert got.dtype == torch.int64
    assert torch.equal(got[clear], ref[clear]), int((got[clear] != ref[clear]).sum())
    assert int((~clear).sum()) <= max(3, b // 1000)


def test_pseudo_labels_exact_ties_pick_lowest_index():
    """Duplicate prototype rows and 4-bit quantised features force exact ties."""
    from latteclip_b200 import prototypes as P
    g = torch.Generator().manual_seed(3)
    c, d, b = 64, 128, 2048
    clf = (torch.randint(-8, 8, (c, d), generator=g).float() / 8.0)
    clf[40] = clf[7]
    clf[63] = clf[7]
    clf[21] = clf[20]
    img = torch.randint(-8, 8, (b, d), generator=g).float() / 8.0     # products exact in fp32
    got = P.pseudo_label(img.to(DEV), clf.to(DEV), 100.0).cpu()
    ref = (100.0 * img @ clf.T).argmax(dim=1)
    assert torch.equal(got, ref)
    assert not bool(((got == 40) | (got == 63) | (got == 21)).any())
    # bf16 / fp16 inputs hold these values exactly as well
    for dt in (torch.bfloat16, torch.float16):
        assert torch.equal(P.pseudo_label(img.to(DEV).to(dt), clf.to(DEV), 100.0).cpu(), ref)


def test_topk_matches_torch():
    from latteclip_b200 import prototypes as P
    bank, _, img, _, _, _ = proto_inputs(3000, 512, 397, 77)
    clf = F.normalize(bank, dim=1)
    idx, val = P.zero_shot_topk(img.to(DEV), clf.to(DEV), 10, 100.0)
    ref = (100.0 * img.double() @ clf.double().T).topk(10, dim=1)
    assert idx.shape == (3000, 10)
    match = (idx.cpu() == ref.indices).float().mean()
    assert float(match) > 0.999
    assert torch.allclose(val.cpu().double(), ref.values, atol=2e-4)
    # top-1 column equals the argmax kernel
    assert torch.equal(idx[:, 0], P.pseudo_label(img.to(DEV), clf.to(DEV), 100.0))


# ------------------------------------------------------------------ K9: margins
def test_text_margins_match_compute_text_weights_golden():
    from latteclip_b200 import prototypes as P
    g = load_golden("text_margins.npz")
    x, p = torch.from_numpy(g["X"]), torch.from_numpy(g["P"])
    got = P.text_margins(x.to(DEV), p.to(DEV)).cpu()
    # cancellation-sensitive: compare at fp32 dot-product accuracy
    assert torch.allclose(got.double(), torch.from_numpy(g["margin_f64"]), atol=5e-7, rtol=0)
    idx = torch.randint(0, x.shape[0], (500,))
    got_g = P.text_margins(x.to(DEV), p.to(DEV), row_index=idx.to(DEV)).cpu()
    assert torch.equal(got_g, got[idx])


# ------------------------------------------------------------------ K10: mixture + EMA
@pytest.mark.parametrize("axis,b,d,c", [("row", 256, 512, 47), ("quirk", 512, 512, 47),
                                         ("row", 1000, 768, 397), ("row", 37, 50, 5)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_mix_and_ema_forward_backward(axis, b, d, c, dtype):
    from latteclip_b200 import prototypes as P
    import oracle
    bank, cls_text, _, pimg, pgrp, zs = proto_inputs(b, d, c, 5 + b)
    g = torch.Generator().manual_seed(b)
    preds = torch.randint(0, c, (b,), generator=g)
    w = [torch.rand(b, generator=g) * 0.3 + 1e-3 for _ in range(4)]
    alpha = 0.01

    def leaf(x):
        return x.to(DEV).to(dtype).requires_grad_(True)
    ct, pi, pg = leaf(cls_text), leaf(pimg), leaf(pgrp)
    t_ft, t_zs = P.mix_and_ema(ct, pi, pg, bank.to(DEV), preds.to(DEV), zs.to(DEV),
                               *[x.to(DEV) for x in w], alpha, axis)
    gf = torch.randn(b, d, generator=g)
    gz = torch.randn(b, d, generator=g)
    (t_ft.float() * gf.to(DEV) + t_zs.float() * gz.to(DEV)).sum().backward()

    def cl(x):
        return x.detach().float().cpu().double().requires_grad_(True)
    ct_c, pi_c, pg_c = cl(ct), cl(pi), cl(pg)
    r_ft, r_zs = oracle.mix_and_ema(ct_c[preds], ct_c[zs], pi_c, pg_c, w[0].double(), w[1].double(),
                                    w[2].double(), w[3].double(), bank.double()[preds],
                                    bank.double()[zs], alpha, axis)
    # the CUDA path receives the upstream gradient rounded to `dtype`
    gf_r = gf.to(dtype).double() if dtype != torch.float32 else gf.double()
    gz_r = gz.to(dtype).double() if dtype != torch.float32 else gz.double()
    (r_ft * gf_r + r_zs * gz_r).sum().backward()
    ftol = 2e-6 if dtype == torch.float32 else 4e-3
    gtol = 1e-5 if dtype == torch.float32 else 8e-3
    assert rel(t_ft, r_ft) < ftol and rel(t_zs, r_zs) < ftol
    assert rel(pi.grad, pi_c.grad) < gtol and rel(pg.grad, pg_c.grad) < gtol
    assert rel(ct.grad, ct_c.grad) < gtol


def test_quirk_axis_requires_square_batch():
    from latteclip_b200 import prototypes as P
    bank, cls_text, _, pimg, pgrp, zs = proto_inputs(64, 32, 5, 1)
    w = torch.rand(64).to(DEV)
    with pytest.raises(RuntimeError, match="must match the size"):
        P.mix_and_ema(cls_text.to(DEV), pimg.to(DEV), pgrp.to(DEV), bank.to(DEV), zs.to(DEV), zs.to(DEV),
                      w, w, w, w, 0.01, "quirk")


# ------------------------------------------------------------------ K11: bank update
@pytest.mark.parametrize("b,d,c", [(512, 512, 47), (8192, 768, 397), (100, 64, 300)])
def test_update_bank(b, d, c):
    from latteclip_b200 import prototypes as P
    import oracle
    bank, _, _, pimg, pgrp, zs = proto_inputs(b, d, c, 11 + b)
    g = torch.Generator().manual_seed(b)
    preds = torch.randint(0, c, (b,), generator=g)
    t_ft, t_zs = pimg * 0.995, pgrp * 0.996
    bank_g = bank.to(DEV).clone()
    _, counts = P.update_bank(bank_g, preds.to(DEV), zs.to(DEV), t_ft.to(DEV), t_zs.to(DEV))
    ref = oracle.update_bank(bank, preds, zs, t_ft, t_zs)
    assert rel(bank_g, ref) < 1e-6
    touched = (torch.bincount(preds, minlength=c) + torch.bincount(zs, minlength=c)) > 0
    assert torch.equal(counts.cpu() > 0, touched)
    # untouched rows are bit-identical
    assert torch.equal(bank_g.cpu()[~touched], bank[~touched])


@pytest.mark.parametrize("b,d,c,dtype", [(32768, 512, 47, torch.float32), (32768, 512, 47, torch.bfloat16),
                                         (20011, 768, 64, torch.float32), (3000, 512, 100, torch.float32),
                                         # > 992 columns: two columns per consumer thread
                                         (1500, 1024, 20, torch.float32), (777, 1536, 12, torch.float16),
                                         # fewer rows than chunks; one class only
                                         (37, 512, 47, torch.float32), (4096, 256, 2, torch.bfloat16)])
def test_streaming_class_sums_at_headline_batch(b, d, c, dtype):
    """The one-pass per-class sums (cls_stream_kernel: [C, D] accumulator in shared memory, fixed row
    chunks) behind latte_bank_accumulate and latte_mix_ema_bwd at the headline batch: values against
    fp64 index_add, exact counts, and bit-identical results from run to run (no atomics)."""
    from latteclip_b200 import _lib
    g = torch.Generator().manual_seed(b + c)
    x = torch.randn(b, d, generator=g).to(DEV).to(dtype)
    y = torch.randn(b, d, generator=g).to(DEV).to(dtype)
    preds = torch.randint(0, c, (b,), generator=g).to(DEV)
    zs = torch.randint(0, c - 1, (b,), generator=g).to(DEV)
    sums, counts = _lib.bank_accumulate(x, y, preds, zs, c)
    ref = torch.zeros(c, d, dtype=torch.float64, device=DEV)
    ref.index_add_(0, zs, y.double())
    ref.index_add_(0, preds, x.double())
    assert rel(sums, ref) < 2e-6
    cnt_ref = torch.bincount(preds, minlength=c) + torch.bincount(zs, minlength=c)
    assert torch.equal(counts.long(), cnt_ref)
    sums2, _ = _lib.bank_accumulate(x, y, preds, zs, c)
    assert torch.equal(sums, sums2)
    # mixer backward: row gradients and class-text sums from one pass
    w = [torch.rand(b, generator=g).to(DEV) * 0.3 + 1e-3 for _ in range(4)]
    alpha = 0.01
    d_ct, d_pi, d_pg, _ = _lib.mix_ema_bwd(x, y, preds, zs, w[0], w[1], w[2], w[3], alpha, "row", c)
    xd, yd = x.double(), y.double()
    wl, wlz, wi, wg = (t.double() for t in w)
    af = alpha * xd / (wl + wi + wg)[:, None]
    az = alpha * yd / (wlz + wi + wg)[:, None]
    ct_ref = torch.zeros(c, d, dtype=torch.float64, device=DEV)
    ct_ref.index_add_(0, preds, wl[:, None] * af)
    ct_ref.index_add_(0, zs, wl[:, None] * az)
    tol = 2e-6 if dtype == torch.float32 else 4e-3      # 16-bit: the row outputs are rounded to dtype
    assert rel(d_ct, ct_ref) < 2e-6
    assert rel(d_pi, (af + az) * wi[:, None]) < tol
    assert rel(d_pg, (af + az) * wg[:, None]) < tol
    d_ct2, d_pi2, _, _ = _lib.mix_ema_bwd(x, y, preds, zs, w[0], w[1], w[2], w[3], alpha, "row", c)
    assert torch.equal(d_ct, d_ct2) and torch.equal(d_pi, d_pi2)


# ------------------------------------------------------------------ full step vs the real train loop
@pytest.mark.parametrize("name", ["b32_c7", "b64_c10", "b64_c10_flags"])
def test_prototype_step_matches_train_one_epoch_v2_golden(name):
    """fp32 features, quirk axis (B == D): every output of our step against what the unmodified
    reference loop produced (tests/golden/make_golden.py)."""
    import latteclip_b200 as lb
    from latteclip_b200 import prototypes as P
    g = load_golden(f"proto_step_{name}.npz")
    nb = int(g["nb"])
    flags = [float(f) for f in g["flags"]]
    scale, alpha = float(g["scale"]), float(g["alpha"])
    bank = torch.from_numpy(g["bank0"]).float().to(DEV).contiguous()
    snapshot = bank.clone()
    loss_fn = lb.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True)
    for b in range(nb):
        def leaf(x):
            return torch.from_numpy(np.asarray(x)).float().to(DEV).requires_grad_(True)
        img, cls_text = leaf(g["img"][b]), leaf(g["cls_text"])
        pimg, pgrp = leaf(g["pimg"][b]), leaf(g["pgrp"][b])
        log_s = torch.tensor(math.log(scale), device=DEV, requires_grad=True)
        zs = torch.from_numpy(g["zs"][b]).to(DEV)
        out = P.prototype_step(img, log_s.exp(), bank, snapshot, zs, cls_text, pimg, pgrp, loss_fn,
                               alpha=alpha, use_image_caption=flags[0], use_batch_caption=flags[1],
                               use_template_caption=flags[2], use_zeroshot_pseudolabel=flags[3],
                               use_finetune_pseudolabel=flags[4], label_weight_axis="quirk")
        out["loss"].backward()
        tol = 2e-5
        assert rel(out["t_ft"], g[f"b{b}_t_ft"]) < tol
        assert rel(out["t_zs"], g[f"b{b}_t_zs"]) < tol
        assert abs(float(out["contrastive_loss"]) - float(g[f"b{b}_loss_ft"])) < 1e-4 * max(1.0, float(g[f"b{b}_loss_ft"]))
        assert abs(float(out["zeroshot"]) - flags[3] * float(g[f"b{b}_loss_zs"])) < 1e-4 * max(1.0, float(g[f"b{b}_loss_zs"]))
        assert rel(img.grad, g[f"b{b}_dI"]) < 1e-4
        assert rel(cls_text.grad, g[f"b{b}_dCls"]) < 1e-4
        assert rel(pimg.grad, g[f"b{b}_dPimg"]) < 1e-4
        if flags[1] != 0.0:
            assert rel(pgrp.grad, g[f"b{b}_dPgrp"]) < 1e-4
        ref_dl = float(g[f"b{b}_dlogscale"])
        assert abs(float(log_s.grad) - ref_dl) < 1e-3 * max(1.0, abs(ref_dl))
        P.update_bank(bank, out["preds"], zs, out["t_ft"], out["t_zs"])
    assert rel(bank, g["final_bank"]) < 1e-5


def test_bank_roundtrip_through_parameter_dict():
    import torch.nn as nn
    from latteclip_b200 import prototypes as P
    names = [f"c{k}" for k in range(5)]
    pd = nn.ParameterDict({n: nn.Parameter(torch.randn(8, device=DEV)) for n in names})
    bank = P.stack_bank(pd, names)
    assert bank.shape == (5, 8)
    bank2 = bank * 2
    P.unstack_bank(bank2, pd, names, touched=torch.tensor([1, 0, 1, 0, 0.0]))
    assert torch.equal(pd["c0"].data, bank2[0]) and torch.equal(pd["c1"].data, bank[1])


# ------------------------------------------------------------------ zero-shot eval
def test_zero_shot_accuracy_matches_reference_golden():
    """zero_shot.py:14-52 / train.py:1128-1138 outputs recorded from the reference."""
    from types import SimpleNamespace
    from latteclip_b200 import zero_shot as zs
    g = load_golden("zero_shot_eval.npz")
    feats, clf = torch.from_numpy(g["feats"]).to(DEV), torch.from_numpy(g["classifier"]).to(DEV)
    target = torch.from_numpy(g["target"]).to(DEV)
    accs, top_logits, top_ids = zs.accuracy(feats, clf, target, topk=(1, 5, 10))
    assert accs == list(g["accs_train"])                        # hit counts: bit-exact
    assert np.array_equal(top_ids.cpu().numpy(), g["top_ids"])  # no ties in this fixture
    assert top_ids.dtype == torch.int64 and top_logits.shape == (feats.shape[0], 10)
    assert np.allclose(top_logits.cpu().numpy(), g["top_logits"], rtol=0, atol=2e-5)

    class Tower(torch.nn.Module):
        def forward(self, image=None):
            return {"image_features": image}

    b = int(g["batch"])
    batches = [(None, feats[k:k + b].cpu(), target[k:k + b].cpu()) for k in range(0, feats.shape[0], b)]
    rates = zs.run(Tower(), clf, batches, SimpleNamespace(precision="fp32", device=DEV))
    assert np.allclose(rates, g["rates"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("b,d,c,dtype", [(5000, 512, 397, torch.bfloat16), (1037, 768, 1000, torch.float32),
                                         (64, 512, 10, torch.float16)])
def test_zero_shot_accuracy_matches_oracle(b, d, c, dtype):
    from oracle import zero_shot as ozs
    from latteclip_b200 import zero_shot as zs
    bank, _, img, _, _, _ = proto_inputs(b, d, c, 5 + b)
    clf = F.normalize(bank, dim=1).T.contiguous()               # [D, C] as zero_shot.py:145
    img = img.to(dtype)
    g = torch.Generator().manual_seed(b)
    target = torch.randint(0, c, (b,), generator=g)
    logits = ozs.zero_shot_logits(img.double(), clf.double())
    want, want_logits, want_ids = ozs.accuracy(logits, target, (1, 5, 10))
    accs, top_logits, top_ids = zs.accuracy(img.to(DEV), clf.to(DEV), target.to(DEV), topk=(1, 5, 10))
    # rows whose 10th/11th (or any adjacent) logits are closer than fp32 dot-product noise may swap
    srt = logits.sort(dim=1, descending=True).values[:, :11]
    tie = ((srt[:, :-1] - srt[:, 1:]).min(dim=1).values <= 2e-4) if c > 10 else torch.zeros(b, dtype=torch.bool)
    ok = ~tie
    assert int(tie.sum()) <= max(2, b // 100)
    assert torch.equal(top_ids.cpu()[ok], want_ids[ok])
    assert torch.allclose(top_logits.cpu().double(), want_logits, rtol=0, atol=2e-4)
    for a, w in zip(accs, want):
        assert abs(a - w) <= int(tie.sum())


def test_extract_feature_records_with_mock_tower(tmp_path):
    """train.py:1336-1381 through the drop-in: a pass-through image tower, records written in the
    reference's pkl format and read back the way data.py:393-396 does."""
    from types import SimpleNamespace
    from latteclip_b200 import zero_shot as zs
    from oracle import zero_shot as ozs
    g = load_golden("zero_shot_eval.npz")
    feats, clf = torch.from_numpy(g["feats"]), torch.from_numpy(g["classifier"]).to(DEV)
    target = torch.from_numpy(g["target"])
    class_names = [f"class {c}" for c in range(clf.shape[1])]
    ids = [f"img_{k:05d}" for k in range(feats.shape[0])]

    class Tower(torch.nn.Module):
        def encode_image(self, images, normalize=True):
            return images

    b = int(g["batch"])
    batches = [(ids[k:k + b], feats[k:k + b], target[k:k + b]) for k in range(0, feats.shape[0], b)]
    args = SimpleNamespace(precision="fp32", device=DEV)
    records, rates = zs.extract_feature_records(Tower(), clf, batches, args, class_names)
    assert np.allclose(rates, g["rates"], rtol=0, atol=1e-12)
    want = ozs.feature_records(ids, feats, torch.from_numpy(g["top_ids"]), torch.from_numpy(g["top_logits"]),
                               target, class_names)
    path = zs.save_feature_records(records, str(tmp_path), "train")
    back = zs.load_key_to_clip_prediction(path)
    assert list(back) == ids
    for k in ids:
        assert np.array_equal(back[k]["top_class_ids"], want[k]["top_class_ids"])
        assert back[k]["class_names"] == want[k]["class_names"]
        assert back[k]["gt_classname"] == want[k]["gt_classname"] and back[k]["gt_class_id"] == want[k]["gt_class_id"]
        assert np.array_equal(back[k]["image"], want[k]["image"])
        assert np.allclose(back[k]["top_logit"], want[k]["top_logit"], rtol=0, atol=2e-5)


def test_prototype_step_and_backward_capture_in_a_cuda_graph():
    """The boundary takes caller workspaces only (no allocation, no host synchronisation inside the
    library), so one whole head step -- pseudo-labels, margins, mixture + EMA, two ClipLoss calls,
    backward, bank update (train.py:384-530) -- can be captured in a CUDA graph and replayed on
    new data; the replay must reproduce the eager result (loss and bank bit for bit)."""
    import latteclip_b200 as lb
    from latteclip_b200 import prototypes as P
    dev = torch.device("cuda:0")
    b = d = 256
    c = 23
    g = torch.Generator().manual_seed(31)

    def make():
        bank = F.normalize(torch.randn(c, d, generator=g), dim=1)
        cls = F.normalize(bank + 0.3 * torch.randn(c, d, generator=g), dim=1)
        true = torch.randint(0, c, (b,), generator=g)
        mk = lambda s: F.normalize(bank[true] + s * torch.randn(b, d, generator=g) * 3 / d ** 0.5, dim=1)
        return dict(bank=bank, cls=cls, img=mk(1.2), pimg=mk(0.9), pgrp=mk(0.7),
                    zs=torch.randint(0, c, (b,), generator=g))

    loss_fn = lb.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True)
    static = {k: v.to(dev).clone() for k, v in make().items()}
    leaves = {k: static[k].bfloat16().requires_grad_(True) for k in ("img", "cls", "pimg", "pgrp")}
    log_s = torch.tensor(math.log(100.0), device=dev, requires_grad=True)
    snap = static["bank"].clone()

    def step():
        out = P.prototype_step(leaves["img"], log_s.exp(), static["bank"], snap, static["zs"], leaves["cls"],
                               leaves["pimg"], leaves["pgrp"], loss_fn, alpha=0.01, label_weight_axis="quirk")
        out["loss"].backward()
        P.update_bank(static["bank"], out["preds"], static["zs"], out["t_ft"], out["t_zs"])
        return out["loss"].detach()

    def load(data):
        with torch.no_grad():
            # the epoch-start snapshot (train.py:347-350) stays what it was at capture time: its
            # operand planes are split once per epoch, outside the captured step
            static["bank"].copy_(data["bank"].to(dev))
            static["zs"].copy_(data["zs"].to(dev))
            for k in leaves:
                leaves[k].copy_(data[k].to(dev).bfloat16())

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):                       # warm-up: workspaces, gradient buffers
            step()
    torch.cuda.current_stream().wait_stream(side)
    for x in list(leaves.values()) + [log_s]:
        x.grad.zero_()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss_static = step()
    static_grads = [(x, x.grad) for x in list(leaves.values()) + [log_s]]   # what the graph writes into
    for trial in range(2):
        data = make()
        load(data)
        for x, gbuf in static_grads:
            x.grad = gbuf
            gbuf.zero_()
        graph.replay()
        torch.cuda.synchronize()
        got = dict(loss=loss_static.clone(), bank=static["bank"].clone(),
                   **{k: leaves[k].grad.clone() for k in leaves}, s=log_s.grad.clone())
        load(data)
        for x in list(leaves.values()) + [log_s]:
            x.grad = None
        want_loss = step()
        torch.cuda.synchronize()
        assert torch.equal(got["loss"], want_loss)
        assert torch.equal(got["bank"], static["bank"])
        # gradients: tiles that the stream-K GEMM splits over more than two clusters are summed with
        # red.global.add in arrival order, so the last fp32 bit (one bf16 ulp after rounding) of each
        # ClipLoss gradient may differ; a leaf that receives two such gradients (autograd adds them in
        # bf16) can cancel, so the bound is one bf16 ulp of the tensor's largest entry per element and
        # 2^-8 of the norm overall, not a per-element relative error
        for k in leaves:
            g_rep, g_eag = got[k].float(), leaves[k].grad.float()
            assert float((g_rep - g_eag).abs().max()) <= 2.0 ** -7 * float(g_eag.abs().max()), k
            assert float((g_rep - g_eag).norm()) <= 2.0 ** -8 * float(g_eag.norm()), k
        assert torch.allclose(got["s"], log_s.grad, rtol=1e-5)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n,d,c", [(1000, 512, 47), (257, 64, 10), (4096, 768, 64), (130, 520, 33),
                                   (20000, 512, 47)])
def test_nxc_multi_one_launch_matches_fp64(dtype, n, d, c):
    """latte_nxc_multi (stacked jobs, in-kernel plane split, persistent TMA pipeline): argmax
    bit-exact wherever the fp64 top-1 / top-2 gap exceeds the fp32 rounding band, margins and top-1
    to fp32 accuracy (train.py:410-411, 292-303); the fused normalisation equals F.normalize."""
    from latteclip_b200 import _lib
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(n + d + c)
    bank = (torch.randn(c, d, generator=g) * 1.7).to(dev)
    snap = F.normalize(torch.randn(c, d, generator=g), dim=1).to(dev) * 0.98
    xs = [F.normalize(torch.randn(m, d, generator=g), dim=1).to(dev).to(dtype) for m in (n, n, max(n // 3, 1), c)]
    cls_planes = _lib.nxc_split_prototypes(bank, normalize=True, want_normalized=True)
    snap_planes = _lib.nxc_split_prototypes(snap)
    assert torch.allclose(cls_planes.normalized, F.normalize(bank, dim=1), rtol=0, atol=1e-7)
    outs = _lib.nxc_multi([dict(x=xs[0], planes=cls_planes, scale=100.0, argmax=True, top1=True),
                           dict(x=xs[1], planes=snap_planes, margin=True),
                           dict(x=xs[2], planes=snap_planes, margin=True, argmax=True),
                           dict(x=xs[3], planes=snap_planes, margin=True)])
    torch.cuda.synchronize()
    refs = [(xs[0].double() @ F.normalize(bank.double(), dim=1).T)] + \
           [x.double() @ snap.double().T for x in xs[1:]]
    # job 0: argmax + scaled top-1
    top = refs[0].topk(2, dim=1)
    gap = top.values[:, 0] - top.values[:, 1]
    clear = gap > 4e-7
    assert torch.equal(outs[0][0][clear], top.indices[:, 0][clear])
    assert int((~clear).sum()) <= max(2, n // 500)
    assert torch.allclose(outs[0][2].double(), 100.0 * top.values[:, 0], rtol=0, atol=3e-5)
    # margins (differences of near-equal dots: absolute tolerance of a few fp32 ulps of a unit dot)
    for k in (1, 2, 3):
        tk = refs[k].topk(2, dim=1).values
        assert torch.allclose(outs[k][1].double(), tk[:, 0] - tk[:, 1], rtol=0, atol=3e-7), k
    tk2 = refs[2].topk(2, dim=1)
    ok = (tk2.values[:, 0] - tk2.values[:, 1]) > 4e-7
    assert torch.equal(outs[2][0][ok], tk2.indices[:, 0][ok])


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_nxc_multi_fp16_planes_on_wide_dynamic_range(dtype):
    """fp32 values enter the tensor core as two fp16 planes, v = h0 + 2^-11 h1 (the low plane stored times
    2^11 so that it stays out of fp16's subnormal range).  Rows whose entries span seven decades and
    prototypes of norm 0.01 .. 30 must still give fp32-accurate dots: the error bound is relative to
    |x| . |p|, as for an fp32 FMA loop (train.py:410-411 runs on whatever the towers emit)."""
    from latteclip_b200 import _lib
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    n, d, c = 3000, 512, 47
    mag = 10.0 ** (-7.0 * torch.rand(n, d, generator=g))              # entries from 1e-7 to 1
    x = (torch.randn(n, d, generator=g) * mag)
    if dtype == torch.float16:
        x = x.clamp(-6e4, 6e4)
    x = x.to(dev).to(dtype)
    protos = torch.randn(c, d, generator=g) * (10.0 ** (torch.rand(c, 1, generator=g) * 3.5 - 2.0)) / d ** 0.5
    protos = protos.to(dev)
    planes = _lib.nxc_split_prototypes(protos)
    (am, mg, t1), = _lib.nxc_multi([dict(x=x, planes=planes, scale=1.0, argmax=True, margin=True, top1=True)])
    ref = x.double() @ protos.double().T
    bound = x.double().norm(dim=1, keepdim=True) * protos.double().norm(dim=1)[None, :]      # |x| |p|
    top = ref.topk(2, dim=1)
    err = (t1.double() - top.values[:, 0]).abs() / bound.gather(1, top.indices[:, :1]).squeeze(1)
    assert float(err.max()) < 2e-7, float(err.max())
    gap = (top.values[:, 0] - top.values[:, 1]) / bound.max(dim=1).values
    clear = gap > 1e-6
    assert torch.equal(am[clear], top.indices[:, 0][clear])
    assert int(clear.sum()) > n // 2


def test_step_similarities_one_launch_equals_per_product_kernels():
    """prototypes.step_similarities through latte_nxc_multi == the per-product kernels it replaces."""
    from latteclip_b200 import prototypes as P
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(17)
    b, d, c = 3000, 512, 47
    bank = F.normalize(torch.randn(c, d, generator=g), dim=1).to(dev)
    snap = bank.clone()
    for dtype in (torch.float32, torch.bfloat16):
        img, pi, pg = (F.normalize(torch.randn(b, d, generator=g), dim=1).to(dev).to(dtype) for _ in range(3))
        ct = F.normalize(torch.randn(c, d, generator=g), dim=1).to(dev).to(dtype)
        preds, m_i, m_g, m_c = P.step_similarities(img, bank, snap, pi, pg, ct)
        ref_preds = P.pseudo_label(img, P.build_classifier(bank), 100.0)
        assert float((preds != ref_preds).float().mean()) < 2e-3
        for got, x in ((m_i, pi), (m_g, pg), (m_c, ct)):
            assert torch.allclose(got, P.text_margins(x, snap), rtol=0, atol=3e-7)
    # mixed operand classes in one step: fp32 images and class texts (converted in the kernel), fp16 and
    # bf16 description texts (MMA operands as they are) -> one launch per class
    img = F.normalize(torch.randn(b, d, generator=g), dim=1).to(dev)
    pi = F.normalize(torch.randn(b, d, generator=g), dim=1).to(dev).half()
    pg = F.normalize(torch.randn(b, d, generator=g), dim=1).to(dev).bfloat16()
    ct = F.normalize(torch.randn(c, d, generator=g), dim=1).to(dev)
    preds, m_i, m_g, m_c = P.step_similarities(img, bank, snap, pi, pg, ct)
    assert float((preds != P.pseudo_label(img, P.build_classifier(bank), 100.0)).float().mean()) < 2e-3
    for got, x in ((m_i, pi), (m_g, pg), (m_c, ct)):
        assert torch.allclose(got, P.text_margins(x, snap), rtol=0, atol=3e-7)


def test_graphed_prototype_step_equals_eager_step():
    """prototypes.GraphedPrototypeStep (one CUDA-graph replay per step, gradients handed to autograd)
    against the eager prototype_step + backward + update_bank on the same data, over three steps with
    a moving bank; upstream gradient 2.5 (as a GradScaler would send) on the last one."""
    import latteclip_b200 as lb
    from latteclip_b200 import prototypes as P
    dev = torch.device("cuda:0")
    b = d = 256
    c = 23
    g = torch.Generator().manual_seed(41)
    bank0 = F.normalize(torch.randn(c, d, generator=g), dim=1)
    cls = F.normalize(bank0 + 0.3 * torch.randn(c, d, generator=g), dim=1)

    def batch():
        true = torch.randint(0, c, (b,), generator=g)
        mk = lambda s: F.normalize(bank0[true] + s * torch.randn(b, d, generator=g) * 3 / d ** 0.5, dim=1)  # noqa: E731
        return dict(img=mk(1.2), pimg=mk(0.9), pgrp=mk(0.7), zs=torch.randint(0, c, (b,), generator=g))

    loss_fn = lb.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True)
    graphed = P.GraphedPrototypeStep(loss_fn, alpha=0.01, label_weight_axis="quirk")
    bank_g, bank_e = bank0.to(dev).clone(), bank0.to(dev).clone()
    snap = bank0.to(dev).clone()
    for it in range(3):
        data = batch()
        up = 2.5 if it == 2 else 1.0
        res = {}
        for tag, bank in (("graph", bank_g), ("eager", bank_e)):
            leaves = {k: data[k].to(dev).bfloat16().requires_grad_(True) for k in ("img", "pimg", "pgrp")}
            ct = cls.to(dev).bfloat16().requires_grad_(True)
            log_s = torch.tensor(math.log(100.0), device=dev, requires_grad=True)
            zs = data["zs"].to(dev)
            if tag == "graph":
                out = graphed(leaves["img"], log_s.exp(), bank, snap, zs, ct, leaves["pimg"], leaves["pgrp"])
                (out["loss"] * up).backward()
            else:
                out = P.prototype_step(leaves["img"], log_s.exp(), bank, snap, zs, ct, leaves["pimg"],
                                       leaves["pgrp"], loss_fn, alpha=0.01, label_weight_axis="quirk")
                (out["loss"] * up).backward()
                P.update_bank(bank, out["preds"], zs, out["t_ft"].detach(), out["t_zs"].detach())
            torch.cuda.synchronize()
            res[tag] = dict(loss=float(out["loss"]), preds=out["preds"].clone(), ct=ct.grad.clone(),
                            s=float(log_s.grad), **{k: v.grad.clone() for k, v in leaves.items()})
        assert abs(res["graph"]["loss"] - res["eager"]["loss"]) < 1e-6 * abs(res["eager"]["loss"])
        assert torch.equal(res["graph"]["preds"], res["eager"]["preds"])
        # bf16 gradients: the two ClipLoss contributions are added and (last step) scaled in bf16, in a
        # different order than autograd's accumulation -- one or two bf16 ulps per entry
        for k in ("img", "pimg", "pgrp", "ct"):
            assert rel(res["graph"][k], res["eager"][k]) < 4e-3, k
        assert abs(res["graph"]["s"] - res["eager"]["s"]) < 1e-4 * abs(res["eager"]["s"]) + 1e-9
        assert torch.allclose(bank_g, bank_e, rtol=0, atol=1e-6)
    assert graphed.graph is not None and graphed.serial == 3
