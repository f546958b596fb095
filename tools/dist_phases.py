"""GPU time of every phase of the peer-memory multi-rank step (CUDA events), torchrun.
Per-kernel-group timing: push | forward (sweep, finalize, payload, merge, finish) | backward phase 1
(prep, sweep, GEMM with the fused reduce-scatter, casts) | backward phase 2 (accumulator finish)."""
import math, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
from latteclip_b200 import _lib
from latteclip_b200 import loss as L
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
N = int(os.environ.get("PHASES_N", "32768"))
n = N // world; D = 512
g = torch.Generator().manual_seed(rank)
img = F.normalize(torch.randn(n, D, generator=g), dim=1).to(dev).bfloat16()
txt = F.normalize(img.float().cpu() + 4.0 * torch.randn(n, D, generator=g) / math.sqrt(D), dim=1).to(dev).bfloat16()
sc = torch.tensor(100.0, device=dev); one = torch.ones(1, device=dev)
state = L._comm_state(n, D, torch.bfloat16, dev, None, world, rank)
assert state is not None, "symmetric memory unavailable"
names = ["push", "fwd_phase1(sweep+finalize+payload)", "fwd_phase2(merge+gated)", "fwd_phase4(finish)",
         "bwd_phase1(prep+sweep+gemm+casts)", "bwd_phase2(acc finish)"]
tot = {k: 0.0 for k in names}
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
off = rank * n
reps = 25
all_marks = []
for it in range(reps):
    slot = state.acquire()
    marks = [ev()]
    _lib.comm_push(slot.comm, txt, None, tensor_stride_bytes=slot.all_img.numel() * 2); marks.append(ev())
    out = None
    for ph in (1, 2, 4):
        res = _lib.clip_fwd_rank(slot.comm, img, slot.all_txt, off, sc, phases=ph, out=out)
        out = res[-1]; marks.append(ev())
    row_all, rown_all, col_all, coln_all, loss, stats, _ = res
    b = None
    for ph in (1, 2):
        b = _lib.clip_bwd(img, txt, None, slot.all_txt, off, sc, row_all, col_all, one, 1.0, True,
                          row_nll_all=rown_all, col_nll_all=coln_all, comm=slot.comm, phases=ph,
                          lse_stats=stats, out=None if b is None else b[-1]); marks.append(ev())
    slot.release(signal=False)
    all_marks.append(marks)
torch.cuda.synchronize()
for marks in all_marks[5:]:
    for k, nm in enumerate(names):
        tot[nm] += marks[k].elapsed_time(marks[k + 1])
step_ms = all_marks[5][0].elapsed_time(all_marks[-1][-1]) / (reps - 5)
t = torch.tensor([tot[k] / (reps - 5) * 1e3 for k in names] + [step_ms * 1e3], device=dev)
allt = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(allt, t)
if rank == 0:
    print(f"world {world} N {N} (us per phase, one line per rank; no host sync inside the loop)")
    print("  rank " + " | ".join(n_[:18] for n_ in names) + " | step")
    for r, v in enumerate(allt):
        print(f"  {r}    " + " | ".join(f"{float(x):18.0f}" for x in v[:-1]) + f" | {float(v[-1]):.0f}", flush=True)
dist.barrier(); dist.destroy_process_group()
