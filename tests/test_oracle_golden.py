"""The oracle (oracle/) against the golden fixtures produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""

import numpy as np
import pytest
import torch

import oracle
from oracle.clip_loss import clip_loss_all_ranks
from conftest import load_golden


def rel(a, b):
    a = torch.as_tensor(np.asarray(a), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


@pytest.mark.parametrize("name", ["small_s100", "small_s14", "ragged_s100", "cfg1_s100"])
def test_clip_w1_matches_reference(name):
    g = load_golden(f"clip_w1_{name}.npz")
    i, t, s = torch.from_numpy(g["I"]), torch.from_numpy(g["T"]), float(g["scale"])
    for tag, dt, tol in (("f64", torch.float64, 1e-12), ("f32", torch.float32, 2e-6)):
        losses, di, dt_, ds = clip_loss_all_ranks([i], [t], s, False, False, dt)
        assert abs(float(losses[0]) - float(g[f"loss_{tag}"])) <= tol * max(1.0, abs(float(g[f"loss_{tag}"])))
        gtol = 1e-6 if tag == "f64" else 1e-5      # golden grads are stored as float32
        assert rel(di[0], g[f"dI_{tag}"]) < gtol
        assert rel(dt_[0], g[f"dT_{tag}"]) < gtol
        # ds is a signed sum over N^2 terms: fp32 summation order alone moves it by ~1e-4 rel
        dstol = 1e-9 if tag == "f64" else 2e-4
        assert abs(float(ds[0]) - float(g[f"ds_{tag}"])) <= max(dstol * abs(float(g[f"ds_{tag}"])), 1e-7 if tag == "f32" else 1e-12)


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("local_loss", [False, True])
@pytest.mark.parametrize("gwg", [False, True])
def test_clip_multirank_emulation_matches_gloo_reference(world, local_loss, gwg):
    g = load_golden(f"clip_dist_w{world}.npz")
    i, t, s = torch.from_numpy(g["I"]), torch.from_numpy(g["T"]), float(g["scale"])
    n = i.shape[0] // world
    ish = [i[r * n:(r + 1) * n] for r in range(world)]
    tsh = [t[r * n:(r + 1) * n] for r in range(world)]
    losses, di, dt_, ds = clip_loss_all_ranks(ish, tsh, s, local_loss, gwg, torch.float64)
    key = f"ll{int(local_loss)}_gwg{int(gwg)}"
    for r in range(world):
        assert abs(float(losses[r]) - float(g[f"{key}_r{r}_loss"])) < 1e-10
        assert rel(di[r], g[f"{key}_r{r}_dI"]) < 1e-10
        assert rel(dt_[r], g[f"{key}_r{r}_dT"]) < 1e-10
        assert abs(float(ds[r]) - float(g[f"{key}_r{r}_ds"])) < 1e-10 * max(1.0, abs(float(g[f"{key}_r{r}_ds"])))


def test_text_margins_match_compute_text_weights():
    g = load_golden("text_margins.npz")
    x, p = torch.from_numpy(g["X"]), torch.from_numpy(g["P"])
    m64 = oracle.text_margins(x.double(), p.double())
    assert torch.allclose(m64, torch.from_numpy(g["margin_f64"]), atol=1e-13, rtol=0)
    m32 = oracle.text_margins(x, p)
    # cancellation-sensitive: absolute tolerance at fp32 dot-product level
    assert torch.allclose(m32, torch.from_numpy(g["margin_f32"]), atol=5e-7, rtol=0)


@pytest.mark.parametrize("name", ["b32_c7", "b64_c10", "b64_c10_flags"])
def test_prototype_step_matches_train_one_epoch_v2(name):
    """Oracle restatement (quirk mode) vs the real train_one_epoch_v2 run through mocks."""
    g = load_golden(f"proto_step_{name}.npz")
    nb = int(g["nb"])
    dt = torch.from_numpy(g["img"]).dtype
    tol = 1e-10 if dt == torch.float64 else 2e-5
    bank = torch.from_numpy(g["bank0"])
    snapshot = bank.clone()                       # train.py:347-350: epoch-start snapshot
    flags = [float(f) for f in g["flags"]]
    scale = float(g["scale"])
    for b in range(nb):
        img = torch.from_numpy(g["img"][b]).clone().requires_grad_(True)
        cls_text = torch.from_numpy(g["cls_text"]).clone().requires_grad_(True)
        pimg = torch.from_numpy(g["pimg"][b]).clone().requires_grad_(True)
        pgrp = torch.from_numpy(g["pgrp"][b]).clone().requires_grad_(True)
        log_s = torch.tensor(np.log(scale), dtype=dt, requires_grad=True)
        zs = torch.from_numpy(g["zs"][b])
        out = oracle.prototype_step(
            img, log_s.exp(), bank, snapshot, zs, cls_text, pimg, pgrp,
            alpha=float(g["alpha"]), use_image_caption=flags[0], use_batch_caption=flags[1],
            use_template_caption=flags[2], use_zeroshot_pseudolabel=flags[3],
            use_finetune_pseudolabel=flags[4], label_weight_axis="quirk")
        out["loss"].backward()
        assert rel(out["t_ft"].detach(), g[f"b{b}_t_ft"]) < tol
        assert rel(out["t_zs"].detach(), g[f"b{b}_t_zs"]) < tol
        assert abs(float(out["contrastive_loss"]) - float(g[f"b{b}_loss_ft"])) < tol * 10
        # the spy records the raw second ClipLoss value (before use_zeroshot_pseudolabel)
        assert abs(float(out["zeroshot"]) - flags[3] * float(g[f"b{b}_loss_zs"])) < tol * 10
        assert rel(img.grad, g[f"b{b}_dI"]) < max(tol, 1e-9) * 10
        assert rel(cls_text.grad, g[f"b{b}_dCls"]) < max(tol, 1e-9) * 10
        assert rel(pimg.grad, g[f"b{b}_dPimg"]) < max(tol, 1e-9) * 10
        if flags[1] != 0.0:
            assert rel(pgrp.grad, g[f"b{b}_dPgrp"]) < max(tol, 1e-9) * 10
        assert abs(float(log_s.grad) - float(g[f"b{b}_dlogscale"])) < max(tol, 1e-9) * 10 * max(1.0, abs(float(g[f"b{b}_dlogscale"])))
        bank = out["bank"]
    assert rel(bank, g["final_bank"]) < max(tol, 1e-12)


def test_row_mode_differs_only_in_label_term():
    """'row' mode == 'quirk' mode when the label weight is constant across samples."""
    torch.manual_seed(0)
    b = d = 16
    x = [torch.randn(b, d, dtype=torch.float64) for _ in range(6)]
    w = torch.full((b,), 0.3, dtype=torch.float64)
    w2, w3, w4 = (torch.rand(b, dtype=torch.float64) + 0.1 for _ in range(3))
    a = oracle.mix_and_ema(x[0], x[1], x[2], x[3], w, w2, w3, w4, x[4], x[5], 0.01, "quirk")
    r = oracle.mix_and_ema(x[0], x[1], x[2], x[3], w, w2, w3, w4, x[4], x[5], 0.01, "row")
    assert torch.allclose(a[0], r[0]) and torch.allclose(a[1], r[1])
    with pytest.raises(RuntimeError):
        oracle.mix_and_ema(x[0][:8], x[1][:8], x[2][:8], x[3][:8], w[:8], w2[:8], w3[:8], w4[:8],
                           x[4][:8], x[5][:8], 0.01, "quirk")


# ------------------------------------------------------------------ zero-shot eval
def test_zero_shot_accuracy_matches_reference_functions():
    from oracle import zero_shot as ozs
    g = load_golden("zero_shot_eval.npz")
    feats, clf = torch.from_numpy(g["feats"]), torch.from_numpy(g["classifier"])
    target = torch.from_numpy(g["target"])
    accs, top_logits, top_ids = ozs.accuracy(ozs.zero_shot_logits(feats, clf), target, (1, 5, 10))
    assert accs == list(g["accs_zero_shot"]) == list(g["accs_train"])      # counts, bit-exact
    assert np.array_equal(top_ids.numpy(), g["top_ids"])
    assert np.array_equal(top_logits.numpy(), g["top_logits"])
    b = int(g["batch"])
    batches = [(feats[k:k + b], target[k:k + b]) for k in range(0, feats.shape[0], b)]
    assert np.allclose(ozs.run_batches(batches, clf), g["rates"], rtol=0, atol=1e-15)


def test_feature_record_format_round_trip(tmp_path):
    """clip_features_{split}.pkl: the product's writer, the reference's reader contract
    (data.py:393-396, 412-416, 448) and the zs-id mapping of train.py:412-417 (host only)."""
    from oracle import zero_shot as ozs
    from latteclip_b200 import zero_shot as zs
    g = load_golden("zero_shot_eval.npz")
    feats, target = torch.from_numpy(g["feats"]), torch.from_numpy(g["target"])
    top_ids, top_logits = torch.from_numpy(g["top_ids"]), torch.from_numpy(g["top_logits"])
    class_names = [f"class {c}" for c in range(int(g["classifier"].shape[1]))]
    image_ids = [f"img_{k:05d}" for k in range(feats.shape[0])]
    want = ozs.feature_records(image_ids, feats, top_ids, top_logits, target, class_names)
    got = zs.feature_records(image_ids, feats, top_ids, top_logits, target, class_names)
    path = zs.save_feature_records(got, str(tmp_path), "train")
    assert path.endswith("clip_features_train.pkl")
    back = zs.load_key_to_clip_prediction(path)
    assert list(back) == image_ids
    for k in image_ids:
        assert set(back[k]) == {"image", "top_class_ids", "class_names", "top_logit", "gt_classname", "gt_class_id"}
        for f in ("image", "top_class_ids", "top_logit"):
            assert np.array_equal(back[k][f], want[k][f]) and back[k][f].dtype == want[k][f].dtype
        assert back[k]["class_names"] == want[k]["class_names"]
        assert back[k]["gt_classname"] == want[k]["gt_classname"]
        assert back[k]["gt_class_id"] == want[k]["gt_class_id"] and isinstance(back[k]["gt_class_id"], int)
    names = [zs.zeroshot_classnames(back[k], 3) for k in image_ids]
    ids = zs.zeroshot_class_ids(names, class_names)
    assert ids.dtype == torch.int64
    assert np.array_equal(ids.numpy(), ozs.zeroshot_class_ids(names, class_names))
    assert np.array_equal(ids.numpy(), g["top_ids"][:, 0])


# ------------------------------------------------------------------ SigLipLoss
@pytest.mark.parametrize("name", ["init", "hot"])
def test_siglip_w1_matches_reference(name):
    from oracle.siglip import siglip_all_ranks
    g = load_golden("siglip.npz")
    i, t = torch.from_numpy(g[f"{name}_I"]), torch.from_numpy(g[f"{name}_T"])
    s, b = float(g[f"{name}_scale"]), float(g[f"{name}_bias"])
    for tag, dt, tol in (("f64", torch.float64, 1e-12), ("f32", torch.float32, 3e-6)):
        lo, di, dt_, ds, db = siglip_all_ranks([i], [t], s, b, dt)
        assert abs(float(lo[0]) - float(g[f"{name}_loss_{tag}"])) <= tol * abs(float(g[f"{name}_loss_{tag}"]))
        gt = 1e-10 if tag == "f64" else 2e-5
        assert rel(di[0], g[f"{name}_dI_{tag}"]) < gt and rel(dt_[0], g[f"{name}_dT_{tag}"]) < gt
        assert abs(float(ds[0]) - float(g[f"{name}_ds_{tag}"])) <= max(gt, 2e-4 if tag == "f32" else 0) * abs(float(g[f"{name}_ds_{tag}"]))
        assert abs(float(db[0]) - float(g[f"{name}_db_{tag}"])) <= max(gt, 2e-4 if tag == "f32" else 0) * abs(float(g[f"{name}_db_{tag}"]))


@pytest.mark.parametrize("world", [2, 3, 4])
def test_siglip_ring_emulation_matches_gloo_reference(world):
    from oracle.siglip import siglip_all_ranks
    g = load_golden("siglip.npz")
    i, t = torch.from_numpy(g[f"w{world}_I"]), torch.from_numpy(g[f"w{world}_T"])
    s, b = float(g[f"w{world}_scale"]), float(g[f"w{world}_bias"])
    n = i.shape[0] // world
    lo, di, dt_, ds, db = siglip_all_ranks([i[r * n:(r + 1) * n] for r in range(world)],
                                           [t[r * n:(r + 1) * n] for r in range(world)], s, b)
    for r in range(world):
        assert abs(float(lo[r]) - float(g[f"w{world}_r{r}_loss"])) <= 1e-12 * abs(float(lo[r]))
        assert rel(di[r], g[f"w{world}_r{r}_dI"]) < 1e-10 and rel(dt_[r], g[f"w{world}_r{r}_dT"]) < 1e-10
        assert abs(float(ds[r]) - float(g[f"w{world}_r{r}_ds"])) <= 1e-10 * abs(float(ds[r]))
        assert abs(float(db[r]) - float(g[f"w{world}_r{r}_db"])) <= 1e-10 * abs(float(db[r]))


@pytest.mark.parametrize("name", ["n96_d64", "n200_d128", "n64_d32_close"])
def test_distill_oracle_matches_reference_class(name):
    """oracle/distill.py against open_clip.loss.DistillClipLoss (tests/golden/distill.npz)."""
    from oracle.distill import distill_clip_loss
    g = load_golden("distill.npz")
    s_s, s_t = (float(v) for v in g[f"{name}_scales"])
    for tag, dt, tol in (("f64", torch.float64, 1e-12), ("f32", torch.float32, 5e-6)):
        il = torch.from_numpy(g[f"{name}_I"]).to(dt).requires_grad_(True)
        tl = torch.from_numpy(g[f"{name}_T"]).to(dt).requires_grad_(True)
        s = torch.tensor(s_s, dtype=dt, requires_grad=True)
        loss = distill_clip_loss(il, tl, s, torch.from_numpy(g[f"{name}_It"]).to(dt),
                                 torch.from_numpy(g[f"{name}_Tt"]).to(dt), torch.tensor(s_t, dtype=dt))
        loss.backward()
        assert abs(float(loss) - float(g[f"{name}_loss_{tag}"])) <= tol * abs(float(g[f"{name}_loss_{tag}"]))
        assert rel(il.grad, g[f"{name}_dI_{tag}"]) < max(tol, 1e-6) * 10
        assert rel(tl.grad, g[f"{name}_dT_{tag}"]) < max(tol, 1e-6) * 10
        assert abs(float(s.grad) - float(g[f"{name}_ds_{tag}"])) <= max(tol * 100, 1e-9) * max(abs(float(g[f"{name}_ds_{tag}"])), 1e-3)
