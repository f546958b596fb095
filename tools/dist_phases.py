"""GPU time of every phase of the one-sweep multi-rank step (CUDA events), torchrun."""
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
n = 32768 // world; D = 512
g = torch.Generator().manual_seed(rank)
img = F.normalize(torch.randn(n, D, generator=g), dim=1).to(dev).bfloat16()
txt = F.normalize(img.float().cpu() + 4.0 * torch.randn(n, D, generator=g) / math.sqrt(D), dim=1).to(dev).bfloat16()
sc = torch.tensor(100.0, device=dev); one = torch.ones(1, device=dev)
slot = L._acquire_gather_slot(n, D, torch.bfloat16, dev, None, world)
peer = L._peer_accumulator(n, D, dev, None)
acc, hdl, ptrs = peer
names = ["barrier0", "push", "barrier1", "fwd_rows", "payload_allgather", "fwd_cols", "acc_zero+barrier", "clip_bwd", "barrier", "cast"]
tot = {k: 0.0 for k in names}
def ev(): 
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
off = rank * n
for it in range(25):
    marks = [ev()]
    slot.hdl.barrier(channel=0); marks.append(ev())
    _lib.push_shards(img, txt, slot.ptrs, rank, slot.buf[0].numel() * 2, slot.multicast_ptr); marks.append(ev())
    slot.hdl.barrier(channel=1); marks.append(ev())
    all_img, all_txt = slot.buf[0], slot.buf[1]
    payload = _lib.clip_fwd_rows(img, all_txt, off, sc); marks.append(ev())
    gathered = L._all_gather_cat(payload.reshape(1, -1), None); marks.append(ev())
    row_all, rown_all, col_all, coln_all, loss, _ = _lib.clip_fwd_cols(gathered, all_img, all_txt, n, off, sc); marks.append(ev())
    acc.zero_(); hdl.barrier(channel=0); marks.append(ev())
    d_img, _, d_s = _lib.clip_bwd(img, txt, all_img, all_txt, off, sc, row_all, col_all, one, 1.0, True,
                                  row_nll_all=rown_all, col_nll_all=coln_all, peer_ptrs=ptrs); marks.append(ev())
    hdl.barrier(channel=1); marks.append(ev())
    d_txt = acc.to(torch.bfloat16); marks.append(ev())
    torch.cuda.synchronize()
    if it >= 5:
        for k, nm in enumerate(names):
            tot[nm] += marks[k].elapsed_time(marks[k + 1])
if rank == 0:
    print(f"world {world}: " + ", ".join(f"{k} {v / 20 * 1e3:.0f}us" for k, v in tot.items()), f"| sum {sum(tot.values()) / 20:.3f} ms", flush=True)
dist.barrier(); dist.destroy_process_group()
