"""Tiny end-to-end case of every kernel family (for compute-sanitizer)."""
import math, os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latteclip_b200 import _lib
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for (n, d) in [(300, 72), (130, 520)]:
    i = F.normalize(torch.randn(n, d, generator=g), dim=1).to(dev).bfloat16()
    t = F.normalize(torch.randn(n, d, generator=g), dim=1).to(dev).bfloat16()
    sc = torch.tensor(30.0, device=dev); one = torch.ones(1, device=dev)
    row, col, loss, rn, cn = _lib.clip_fwd(i, t, i, t, 0, sc, with_nll=True)
    di, dt, ds = _lib.clip_bwd(i, t, i, t, 0, sc, row, col, one, 1.0, True, row_nll_all=rn, col_nll_all=cn)
    # two emulated ranks: one-sweep flow
    h = n // 2
    gathered = torch.stack([_lib.clip_fwd_rows(i[r * h:(r + 1) * h], t[:2 * h], r * h, sc) for r in range(2)])
    ra, rna, ca, cna, l0 = _lib.clip_fwd_cols(gathered, i[:2 * h], t[:2 * h], h, 0, sc)
    d0, dp, dss = _lib.clip_bwd(i[:h], t[:h], i[:2 * h], t[:2 * h], 0, sc, ra, ca, one, 1.0, True,
                                row_nll_all=rna, col_nll_all=cna, partial=True)
    print(n, d, float(loss), float(l0), float(di.float().norm()), float(dp.norm()))
c, d, b = 47, 72, 300
bank = F.normalize(torch.randn(c, d, generator=g), dim=1).to(dev)
x = F.normalize(torch.randn(b, d, generator=g), dim=1).to(dev)
p = F.normalize(torch.randn(b, d, generator=g), dim=1).to(dev)
am, mg, _ = _lib.nxc_argmax_margin(x, bank, scale=100.0)
idx, val = _lib.nxc_topk(x, bank, 5, scale=100.0)
preds = am; zs = torch.randint(0, c, (b,), generator=g).to(dev)
w = [torch.rand(b, generator=g).to(dev) + 0.1 for _ in range(4)]
tf, tz = _lib.mix_ema_fwd(bank, x, p, bank, preds, zs, w[0], w[1], w[2], w[3], 0.01, "row")
_lib.mix_ema_bwd(tf, tz, preds, zs, w[0], w[1], w[2], w[3], 0.01, "row", c)
sums, counts = _lib.bank_accumulate(tf, tz, preds, zs, c)
_lib.bank_finalize(sums, counts, bank)
torch.cuda.synchronize()
print("ok", float(bank.norm()), int(counts.sum()))
