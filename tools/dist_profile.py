"""torch.profiler view of one rank's multi-rank ClipLoss step (launch with torchrun)."""
import math, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
import latteclip_b200 as lb
from torch.profiler import profile, ProfilerActivity
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
n = 32768 // world
g = torch.Generator().manual_seed(rank)
i = F.normalize(torch.randn(n, 512, generator=g), dim=1).to(dev).bfloat16().requires_grad_(True)
t = F.normalize(torch.randn(n, 512, generator=g), dim=1).to(dev).bfloat16().requires_grad_(True)
log_s = torch.tensor(math.log(100.0), device=dev, requires_grad=True)
fn = lb.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True, rank=rank, world_size=world)
def step():
    i.grad = None; t.grad = None; log_s.grad = None
    loss = fn(i, t, log_s.exp()); loss.backward()
for _ in range(10): step()
torch.cuda.synchronize(); dist.barrier()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(10): step()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start; t1 = max(e.time_range.end for e in evs)
    busy = 0; cur_end = t0
    for e in evs:
        s_, e_ = e.time_range.start, e.time_range.end
        if e_ > cur_end:
            busy += e_ - max(s_, cur_end); cur_end = e_
    print(f"wall {(t1-t0)/10:.1f} us/step, GPU busy (union) {busy/10:.1f} us/step")
    from collections import defaultdict
    agg = defaultdict(float)
    for e in evs: agg[e.name[:60]] += (e.time_range.end - e.time_range.start) / 10
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:22]: print(f"{v:9.1f} us  {k}")
dist.barrier(); dist.destroy_process_group()
