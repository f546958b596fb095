"""Host-side enqueue cost of one ClipLoss fwd+bwd (no device sync inside the loop) vs device time.
Launch with torchrun for N > 1."""
import math, os, sys, time
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
import latteclip_b200 as lb
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
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
torch.cuda.synchronize()
if world > 1: dist.barrier()
K = 50
t0 = time.perf_counter()
for _ in range(K): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
if rank == 0:
    print(f"world {world}: host enqueue {1e3*(t1-t0)/K:.3f} ms/step, total {1e3*(t2-t0)/K:.3f} ms/step")
if world > 1: dist.destroy_process_group()
