"""Achieved HBM GB/s of the prototype-path kernels (CUDA events, warm, rotating buffers > L2)."""
import json, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latteclip_b200 import _lib

dev = torch.device("cuda:0")
PEAK = 6549.8
if os.path.exists("MEASURED_PEAKS.json"):
    PEAK = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])


def timeit(fn, reps=20):
    """ms per call from a CUDA-graph replay of `reps` captured calls (host overhead excluded)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for k in range(3):
            fn(k)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for k in range(reps):
            fn(k)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(B=32768, D=512, C=47, dtype=torch.float32, sets=6):
    g = torch.Generator().manual_seed(0)
    b = 4 if dtype == torch.float32 else 2
    bank = F.normalize(torch.randn(C, D, generator=g), dim=1).to(dev)
    cls = F.normalize(torch.randn(C, D, generator=g), dim=1).to(dev).to(dtype)
    xs = [F.normalize(torch.randn(B, D, generator=g), dim=1).to(dev).to(dtype) for _ in range(sets)]
    ps = [F.normalize(torch.randn(B, D, generator=g), dim=1).to(dev).to(dtype) for _ in range(sets)]
    preds = torch.randint(0, C, (B,), generator=g).to(dev)
    zs = torch.randint(0, C, (B,), generator=g).to(dev)
    w = [torch.rand(B, generator=g).to(dev) + 0.1 for _ in range(4)]
    out = []
    t = timeit(lambda k: _lib.nxc_argmax_margin(xs[k % sets], bank, scale=100.0, want_argmax=True, want_margin=False))
    out.append(("nxc argmax", B * D * b + C * D * 4 + B * 8, t))
    t = timeit(lambda k: _lib.nxc_argmax_margin(xs[k % sets], bank, scale=1.0, want_argmax=False, want_margin=True))
    out.append(("nxc margin", B * D * b + C * D * 4 + B * 4, t))
    t = timeit(lambda k: _lib.nxc_topk(xs[k % sets], bank, 10, scale=100.0))
    out.append(("nxc top-10", B * D * b + C * D * 4 + B * 10 * 12, t))
    t = timeit(lambda k: _lib.mix_ema_fwd(cls, xs[k % sets], ps[k % sets], bank, preds, zs, w[0], w[1], w[2], w[3], 0.01, "row"))
    out.append(("mix_ema fwd", 4 * B * D * b + 2 * C * D * 4 + 6 * B * 4 + 2 * B * 8, t))
    tf, tz = _lib.mix_ema_fwd(cls, xs[0], ps[0], bank, preds, zs, w[0], w[1], w[2], w[3], 0.01, "row")
    t = timeit(lambda k: _lib.mix_ema_bwd(xs[k % sets], ps[k % sets], preds, zs, w[0], w[1], w[2], w[3], 0.01, "row", C))
    out.append(("mix_ema bwd", 2 * B * D * b + 2 * B * D * b + C * D * 4 + 6 * B * 4, t))
    t = timeit(lambda k: _lib.bank_accumulate(xs[k % sets], ps[k % sets], preds, zs, C))
    out.append(("bank accumulate", 2 * B * D * b + 2 * B * 8 + C * D * 4, t))
    print(f"B={B} D={D} C={C} {dtype}:")
    for name, byts, ms in out:
        gbs = byts / (ms * 1e-3) / 1e9
        print(f"  {name:16s} {ms*1e3:8.1f} us  {byts/1e6:8.1f} MB  {gbs:7.0f} GB/s  {gbs/PEAK:5.2f} of measured HBM peak")


for C in (47, 397, 1000):
    run(C=C, dtype=torch.float32)
run(C=47, dtype=torch.bfloat16)
run(B=512, C=47, dtype=torch.float32)
