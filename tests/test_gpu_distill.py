"""DistillClipLoss (second half) on the GPU through the C ABI: against the golden
vectors recorded from open_clip.loss.DistillClipLoss and against the fp64 oracle on the same rounded
operands at tensor-core sizes."""

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a = a.detach().double().cpu() if torch.is_tensor(a) else torch.as_tensor(np.asarray(a), dtype=torch.float64)
    b = b.detach().double().cpu() if torch.is_tensor(b) else torch.as_tensor(np.asarray(b), dtype=torch.float64)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def pairs(n, d, sigma, seed):
    g = torch.Generator().manual_seed(seed)
    i = F.normalize(torch.randn(n, d, generator=g), dim=1)
    t = F.normalize(i + sigma * torch.randn(n, d, generator=g) / d ** 0.5, dim=1)
    return i, t


@pytest.mark.parametrize("name", ["n96_d64", "n200_d128", "n64_d32_close"])
def test_distill_matches_reference_golden(name):
    """fp32 features (as the golden's) -> fp16 tensor-core operands: the loss and the gradients of the
    reference class within the fp16-operand tolerance (2^-11 per feature, 2^-12 per softmax weight)."""
    import latteclip_b200 as lb
    g = load_golden("distill.npz")
    s_s, s_t = (float(v) for v in g[f"{name}_scales"])
    il = torch.from_numpy(g[f"{name}_I"]).float().to(DEV).requires_grad_(True)
    tl = torch.from_numpy(g[f"{name}_T"]).float().to(DEV).requires_grad_(True)
    it = torch.from_numpy(g[f"{name}_It"]).float().to(DEV)
    tt = torch.from_numpy(g[f"{name}_Tt"]).float().to(DEV)
    s = torch.tensor(s_s, device=DEV, requires_grad=True)
    mod = lb.DistillClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True)
    out = mod(il, tl, s, it, tt, torch.tensor(s_t, device=DEV), output_dict=True)
    assert out["contrastive_loss"] == 0 and set(out) == {"contrastive_loss", "distill_loss"}
    zero, loss = mod(il, tl, s, it, tt, s_t)                    # positional form, python-float teacher scale
    assert zero == 0
    loss.backward()
    ref = float(g[f"{name}_loss_f64"])
    assert abs(float(out["distill_loss"]) - ref) < 3e-3 * abs(ref)
    assert abs(float(loss) - ref) < 3e-3 * abs(ref)
    close = name.endswith("close")
    # "close": student ~ teacher, the gradient is the small difference of two O(1) softmax weights
    gtol = 0.06 if close else 6e-3
    assert rel(il.grad, g[f"{name}_dI_f64"]) < gtol
    assert rel(tl.grad, g[f"{name}_dT_f64"]) < gtol
    ds_ref = float(g[f"{name}_ds_f64"])
    assert abs(float(s.grad) - ds_ref) < (0.1 if close else 1e-2) * abs(ds_ref) + 1e-6


@pytest.mark.parametrize("n,d,dtype", [(1024, 512, torch.bfloat16), (3000, 512, torch.float16),
                                        (2048, 768, torch.bfloat16), (4096, 256, torch.bfloat16)])
def test_distill_matches_fp64_oracle_on_rounded_operands(n, d, dtype):
    """16-bit features: the CUDA path against oracle/distill.py in fp64 on the same rounded features
    (tolerance of north_star for 16-bit inputs: gradients 2e-3 relative, widened to 3e-3 for the
    second fp16 rounding of the softmax weights)."""
    import latteclip_b200 as lb
    from oracle.distill import distill_clip_loss
    i_s, t_s = pairs(n, d, 3.0, n + d)
    i_t, t_t = pairs(n, d, 1.5, n + d + 1)
    s_s, s_t = 30.0, 100.0
    il = i_s.to(DEV).to(dtype).requires_grad_(True)
    tl = t_s.to(DEV).to(dtype).requires_grad_(True)
    it, tt = i_t.to(DEV).to(dtype), t_t.to(DEV).to(dtype)
    log_s = torch.tensor(float(np.log(s_s)), device=DEV, requires_grad=True)
    mod = lb.DistillClipLoss()
    _, loss = mod(il, tl, log_s.exp(), it, tt, torch.tensor(s_t, device=DEV))
    loss.backward()

    def c(x):
        return x.detach().float().cpu().double()
    ic, tc = c(il).requires_grad_(True), c(tl).requires_grad_(True)
    sc = torch.tensor(s_s, dtype=torch.float64, requires_grad=True)
    ref = distill_clip_loss(ic, tc, sc, c(it), c(tt), torch.tensor(s_t, dtype=torch.float64))
    ref.backward()
    assert abs(float(loss) - float(ref)) < 5e-4 * abs(float(ref))
    # the gradients come back rounded to the 16-bit feature dtype
    gtol = 3e-3 if dtype == torch.float16 else 4e-3
    assert rel(il.grad, ic.grad) < gtol
    assert rel(tl.grad, tc.grad) < gtol
    assert abs(float(log_s.grad) / s_s - float(sc.grad)) < 5e-3 * abs(float(sc.grad)) + 1e-7


def test_distill_gradient_vanishes_for_identical_models():
    """Student == teacher: W(S) - W(S') = 0 (same kernels, same operands; the stream-K GEMM adds its
    fp32 partial tiles in a run-dependent order, so the two products agree to fp32 rounding, not
    bit for bit): the feature gradients vanish and the loss is the mean entropy of the softmaxes."""
    import latteclip_b200 as lb
    i_s, t_s = pairs(512, 256, 2.0, 9)
    il = i_s.to(DEV).bfloat16().requires_grad_(True)
    tl = t_s.to(DEV).bfloat16().requires_grad_(True)
    s = torch.tensor(50.0, device=DEV)
    _, loss = lb.DistillClipLoss()(il, tl, s, il.detach(), tl.detach(), s)
    loss.backward()
    assert float(il.grad.abs().max()) < 1e-7 and float(tl.grad.abs().max()) < 1e-7      # W.T entries are O(1)
    logits = 50.0 * il.detach().double() @ tl.detach().double().T
    ent = 0.5 * (-(logits.softmax(1) * logits.log_softmax(1)).sum(1).mean()
                 - (logits.T.softmax(1) * logits.T.log_softmax(1)).sum(1).mean())
    # the loss is the small difference of two terms of size s = 50 (sum of LSEs and s <I, W T>), W in fp16:
    # absolute error ~ 1e-5 * s
    assert abs(float(loss) - float(ent)) < 2e-5 * 50.0
