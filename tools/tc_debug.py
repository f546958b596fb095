"""Diagnostic (not a test): run the fused ClipLoss kernels on a few shapes and print error
magnitudes against plain torch on the same GPU.  Usage: python tools/tc_debug.py [dtype]"""
import math
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latteclip_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
dtype = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[sys.argv[1] if len(sys.argv) > 1 else "bf16"]


def synth(n, d, sigma, seed):
    g = torch.Generator().manual_seed(seed)
    i = F.normalize(torch.randn(n, d, generator=g), dim=1)
    t = F.normalize(i + sigma * torch.randn(n, d, generator=g) / math.sqrt(d), dim=1)
    return i.to(dev).to(dtype), t.to(dev).to(dtype)


def torch_ref(i, t, scale):
    il = i.float().requires_grad_(True)
    tl = t.float().requires_grad_(True)
    s = torch.tensor(scale, device=dev, requires_grad=True)
    li = (s * il) @ tl.T
    lt = (s * tl) @ il.T
    lab = torch.arange(i.shape[0], device=dev)
    row_lse = torch.logsumexp(li, 1)
    col_lse = torch.logsumexp(lt, 1)
    loss = (F.cross_entropy(li, lab) + F.cross_entropy(lt, lab)) / 2
    loss.backward()
    return row_lse.detach(), col_lse.detach(), loss.detach(), il.grad, tl.grad, s.grad


for (n, d, scale) in [(128, 64, 100.0), (256, 512, 100.0), (200, 200, 100.0), (1024, 512, 14.28),
                      (1000, 768, 100.0), (4096, 512, 100.0)]:
    i, t = synth(n, d, 4.0, n + d)
    sc = torch.tensor(scale, device=dev)
    try:
        row, col, loss = _lib.clip_fwd(i, t, i, t, 0, sc)
        torch.cuda.synchronize()
        r_row, r_col, r_loss, r_di, r_dt, r_ds = torch_ref(i, t, scale)
        print(f"[{n}x{d} s={scale}] fwd: row_lse err {float((row - r_row).abs().max()):.3e} "
              f"col_lse err {float((col - r_col).abs().max()):.3e} loss {float(loss):.6f} ref {float(r_loss):.6f}",
              flush=True)
        di, dt, ds = _lib.clip_bwd(i, t, i, t, 0, sc, row, col, torch.ones(1, device=dev), 1.0, True)
        torch.cuda.synchronize()
        e_i = float((di.float() - r_di).norm() / r_di.norm())
        e_t = float((dt.float() - r_dt).norm() / r_dt.norm())
        print(f"[{n}x{d}] bwd: dI rel {e_i:.3e} dT rel {e_t:.3e} ds {float(ds):.6e} ref {float(r_ds):.6e}",
              flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"[{n}x{d}] FAILED: {e}", flush=True)
        break

# timing at the headline size
try:
    n, d = 32768, 512
    i, t = synth(n, d, 4.0, 1)
    sc = torch.tensor(100.0, device=dev)
    for _ in range(2):
        row, col, loss = _lib.clip_fwd(i, t, i, t, 0, sc)
        di, dt, ds = _lib.clip_bwd(i, t, i, t, 0, sc, row, col, torch.ones(1, device=dev), 1.0, True)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    row, col, loss = _lib.clip_fwd(i, t, i, t, 0, sc)
    e1.record()
    di, dt, ds = _lib.clip_bwd(i, t, i, t, 0, sc, row, col, torch.ones(1, device=dev), 1.0, True)
    e2.record()
    torch.cuda.synchronize()
    tf, tb = e0.elapsed_time(e1), e1.elapsed_time(e2)
    fl = 2.0 * n * n * d
    print(f"[32768x512 {dtype}] fwd {tf:.3f} ms ({2 * fl / tf / 1e9:.1f} TF/s executed)  "
          f"bwd {tb:.3f} ms  total {tf + tb:.3f} ms -> {n / (tf + tb) * 1e3 / 1e6:.3f} M samples/s, "
          f"credited {3 * fl / (tf + tb) / 1e9:.1f} TF/s  loss {float(loss):.5f}", flush=True)
except Exception as e:  # noqa: BLE001
    print("timing FAILED:", e, flush=True)
