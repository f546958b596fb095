"""One warm-up and one measured fused ClipLoss fwd+bwd at the headline size, for ncu."""
import math, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latteclip_b200 import _lib
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
g = torch.Generator().manual_seed(1)
i = F.normalize(torch.randn(n, d, generator=g), dim=1)
t = F.normalize(i + 4.0 * torch.randn(n, d, generator=g) / math.sqrt(d), dim=1)
i, t = i.to(dev).bfloat16(), t.to(dev).bfloat16()
sc = torch.tensor(100.0, device=dev)
one = torch.ones(1, device=dev)
for _ in range(2):
    row, col, loss = _lib.clip_fwd(i, t, i, t, 0, sc)
    di, dt, ds = _lib.clip_bwd(i, t, i, t, 0, sc, row, col, one, 1.0, True)
torch.cuda.synchronize()
print("loss", float(loss), "ds", float(ds))
