"""Single-GPU emulation of W ranks (tests/test_gpu_clip.py:_EmulatedRanks): per-phase kernel time of
the peer-memory flow WITHOUT any NVLink traffic or cross-rank waiting -- what the kernels cost by
themselves.  usage: python tools/emul_phase_times.py [world] [n_per_rank]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from latteclip_b200 import _lib
from test_gpu_clip import _EmulatedRanks, synth
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
d = 512
dev = torch.device("cuda:0")
em = _EmulatedRanks(world, n, d, torch.float16, dev)
i_all, t_all = synth(n * world, d, 4.0, 5)
ib, tb = i_all.to(dev).half(), t_all.to(dev).half()
sc = torch.tensor(100.0, device=dev); one = torch.ones(1, device=dev)
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
names = ["push", "fwd1", "fwd2", "fwd4", "bwd1", "bwd2"]
tot = {k: 0.0 for k in names}
reps = 6
for gen in range(1, reps + 1):
    em.set_gen(gen)
    marks = {k: [] for k in names}
    for r in range(world):
        a = ev(); _lib.comm_push(em.comms[r], tb[r * n:(r + 1) * n], None, tensor_stride_bytes=em.N * d * 2); marks["push"].append((a, ev()))
    outs = [None] * world
    for ph, nm in ((1, "fwd1"), (2, "fwd2"), (4, "fwd4")):
        for r in range(world):
            a = ev()
            res = _lib.clip_fwd_rank(em.comms[r], ib[r * n:(r + 1) * n], em.gather[r][1], r * n, sc, phases=ph, out=outs[r])
            marks[nm].append((a, ev()))
            outs[r] = res if ph == 4 else res[-1]
    bwd = [None] * world
    for ph, nm in ((1, "bwd1"), (2, "bwd2")):
        for r in range(world):
            row_all, rown_all, col_all, coln_all, loss_r, stats, _ = outs[r]
            sl = slice(r * n, (r + 1) * n)
            a = ev()
            bwd[r] = _lib.clip_bwd(ib[sl], tb[sl], None, em.gather[r][1], r * n, sc, row_all, col_all, one, 1.0, True,
                                   row_nll_all=rown_all, col_nll_all=coln_all, comm=em.comms[r], phases=ph,
                                   lse_stats=stats, out=None if bwd[r] is None else bwd[r][-1])
            marks[nm].append((a, ev()))
    torch.cuda.synchronize()
    if gen > 2:
        for k in names:
            tot[k] += sum(a.elapsed_time(b) for a, b in marks[k]) / world
print(f"emulated world {world}, n {n}: " + ", ".join(f"{k} {tot[k] / (reps - 2) * 1e3:.0f}us" for k in names))
