"""Eager calls of the class-sum / mixer-backward / stacked N x C kernels at 32768 x 512 x 47 -- the target
of the `ncu --set full` captures of the prototype path (no CUDA graph, a few calls each)."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latteclip_b200 import _lib

dev = torch.device("cuda:0")
B, D, C = 32768, 512, 47
g = torch.Generator().manual_seed(0)
bank = F.normalize(torch.randn(C, D, generator=g), dim=1).to(dev)
preds = torch.randint(0, C, (B,), generator=g).to(dev)
zs = torch.randint(0, C, (B,), generator=g).to(dev)
w = [torch.rand(B, generator=g).to(dev) + 0.1 for _ in range(4)]
planes = _lib.nxc_split_prototypes(bank)
for dtype in (torch.float32, torch.bfloat16):
    xs = [F.normalize(torch.randn(B, D, generator=g), dim=1).to(dev).to(dtype) for _ in range(4)]
    for k in range(3):
        _lib.bank_accumulate(xs[k], xs[(k + 1) % 4], preds, zs, C)
        _lib.mix_ema_bwd(xs[k], xs[(k + 1) % 4], preds, zs, w[0], w[1], w[2], w[3], 0.01, "row", C)
        _lib.nxc_multi([dict(x=xs[(k + j) % 4], planes=planes, scale=100.0, argmax=(j == 0), margin=(j > 0))
                        for j in range(4)])
torch.cuda.synchronize()
print("ok")
