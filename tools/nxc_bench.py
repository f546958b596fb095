"""HBM GB/s of the one-launch N x C path (latte_nxc_multi) beside the per-product kernels."""
import json, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latteclip_b200 import _lib
dev = torch.device("cuda:0")
PEAK = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]) if os.path.exists("MEASURED_PEAKS.json") else 6549.8

def timeit(fn, reps=20):
    """ms per call: `reps` calls captured in ONE CUDA graph and replayed, so the host side of the call
    (ctypes, tensor-map encoding, ~20 us) is not what is measured."""
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
    xs = [F.normalize(torch.randn(B, D, generator=g), dim=1).to(dev).to(dtype) for _ in range(sets)]
    planes = _lib.nxc_split_prototypes(bank)
    t_split = timeit(lambda k: _lib.nxc_split_prototypes(bank, normalize=True))
    t1 = timeit(lambda k: _lib.nxc_multi([dict(x=xs[k % sets], planes=planes, scale=100.0, argmax=True)]))
    t3 = timeit(lambda k: _lib.nxc_multi([dict(x=xs[k % sets], planes=planes, scale=100.0, argmax=True),
                                          dict(x=xs[(k + 1) % sets], planes=planes, margin=True),
                                          dict(x=xs[(k + 2) % sets], planes=planes, margin=True)]))
    told = timeit(lambda k: _lib.nxc_argmax_margin(xs[k % sets], bank, scale=100.0, want_argmax=True, want_margin=False))
    by1 = B * D * b + B * 8
    print(f"B={B} D={D} C={C} {dtype}: split {t_split*1e3:.1f}us | one job {t1*1e3:.1f}us {by1/t1/1e6:.0f} GB/s "
          f"({by1/t1/1e6/PEAK:.2f} of HBM) | three jobs {t3*1e3:.1f}us {(3*B*D*b+B*16)/t3/1e6:.0f} GB/s "
          f"({(3*B*D*b+B*16)/t3/1e6/PEAK:.2f}) | per-product kernel {told*1e3:.1f}us ({by1/told/1e6/PEAK:.2f})", flush=True)

for dt in (torch.float32, torch.bfloat16, torch.float16):
    run(dtype=dt)
run(B=512, dtype=torch.float32)
run(B=512, dtype=torch.bfloat16)
run(C=64, D=768, dtype=torch.float32)
