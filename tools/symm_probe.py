"""Probe: does torch symmetric memory (peer-mapped buffers) work on this box?  torchrun, 2 ranks."""
import os, sys, torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as sm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
try:
    t = sm.empty(1024, 512, dtype=torch.float32, device=dev)
    t.zero_()
    hdl = sm.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok; ptrs", [hex(p) for p in hdl.buffer_ptrs], flush=True)
    hdl.barrier(channel=0)
    peer = (rank + 1) % world
    pt = hdl.get_buffer(peer, (1024, 512), torch.float32)
    pt[rank].add_(float(rank + 1))          # plain peer store through the mapping
    torch.cuda.synchronize()
    hdl.barrier(channel=0)
    torch.cuda.synchronize()
    src = (rank - 1) % world
    print(rank, "row written by peer:", float(t[src, 0]), "expected", float(src + 1), flush=True)
except Exception as e:
    print(rank, "FAILED:", repr(e)[:500], flush=True)
dist.barrier(); dist.destroy_process_group()
