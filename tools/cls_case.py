"""bank_accumulate at the headline shape a few times (ncu target for cls_stream_kernel)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latteclip_b200 import _lib
dev = torch.device("cuda:0")
B, D, C = 32768, 512, 47
g = torch.Generator().manual_seed(0)
dt = torch.bfloat16 if len(sys.argv) > 1 and sys.argv[1] == "bf16" else torch.float32
xs = [torch.randn(B, D, generator=g).to(dev).to(dt) for _ in range(4)]
preds = torch.randint(0, C, (B,), generator=g).to(dev)
zs = torch.randint(0, C, (B,), generator=g).to(dev)
for k in range(6):
    _lib.bank_accumulate(xs[k % 4], xs[(k + 1) % 4], preds, zs, C)
torch.cuda.synchronize()
print("done")
