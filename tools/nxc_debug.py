import os, sys, torch, numpy as np
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latteclip_b200 import _lib
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for (n, c, d) in [(4096, 47, 512), (4096, 397, 768), (100, 47, 512)]:
    x = F.normalize(torch.randn(n, d, generator=g), dim=1)
    p = F.normalize(torch.randn(c, d, generator=g), dim=1)
    ref = (x.double() @ p.double().T)
    r1 = ref.max(1).values
    t2 = ref.topk(2, dim=1).values
    am, mg, t1 = _lib.nxc_argmax_margin(x.to(dev), p.to(dev), scale=1.0, want_argmax=True, want_margin=True, want_top1=True)
    e1 = (t1.cpu().double() - r1)
    em = (mg.cpu().double() - (t2[:, 0] - t2[:, 1]))
    f32 = (x @ p.T).double()
    ef = f32.max(1).values - r1
    print(f"n={n} c={c} d={d}: top1 err mean {float(e1.mean()):.3e} max|.| {float(e1.abs().max()):.3e} | margin err max {float(em.abs().max()):.3e} | torch fp32 matmul top1 err max {float(ef.abs().max()):.3e} | argmax mismatches {(am.cpu() != ref.argmax(1)).sum().item()}")
gd = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/text_margins.npz"))
x, p = torch.from_numpy(gd["X"]), torch.from_numpy(gd["P"])
print("golden shapes", x.shape, p.shape, "norms", float(x.norm(dim=1).mean()), float(p.norm(dim=1).mean()))
_, mg, _ = _lib.nxc_argmax_margin(x.to(dev), p.to(dev), scale=1.0, want_argmax=False, want_margin=True)
print("golden margin err max", float((mg.cpu().double() - torch.from_numpy(gd["margin_f64"])).abs().max()))
