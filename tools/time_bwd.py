"""Time clip_fwd / clip_bwd at the headline size (CUDA events, several reps)."""
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
row, col, loss = _lib.clip_fwd(i, t, i, t, 0, sc)
for _ in range(3):
    _lib.clip_bwd(i, t, i, t, 0, sc, row, col, one, 1.0, True)
    _lib.clip_fwd(i, t, i, t, 0, sc)
torch.cuda.synchronize()
reps = 10
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
for _ in range(reps):
    _lib.clip_fwd(i, t, i, t, 0, sc)
e[1].record()
for _ in range(reps):
    _lib.clip_bwd(i, t, i, t, 0, sc, row, col, one, 1.0, True)
e[2].record()
torch.cuda.synchronize()
print(f"n={n} d={d}: fwd {e[0].elapsed_time(e[1])/reps:.3f} ms  bwd {e[1].elapsed_time(e[2])/reps:.3f} ms")
