"""SigLipLoss fwd+bwd at global batch 32768, dim 512, bf16 on N ranks (strong scaling), timed like
bench.py (barrier + CUDA events, max over ranks):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/siglip_dist_bench.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import latteclip_b200 as lb  # noqa: E402

if __name__ == "__main__":
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_glob, dim, steps = 32768, 512, 20
    sets = []
    for k in range(4):
        i, t = bench.synth_shard(n_glob, dim, rank, world, set_id=k)
        sets.append((i.to(dev).bfloat16(), t.to(dev).bfloat16()))
    s = torch.tensor(10.0, device=dev, requires_grad=True)
    b = torch.tensor(-10.0, device=dev, requires_grad=True)
    mod = lb.SigLipLoss(rank=rank, world_size=world)

    def step(k):
        i, t = sets[k % 4]
        i = i.detach().requires_grad_(True)
        t = t.detach().requires_grad_(True)
        s.grad = b.grad = None
        loss = mod(i, t, s, b)
        loss.backward()
        return loss

    for k in range(5):
        step(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        loss = step(k)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"workload": "SigLipLoss fwd+bwd, global batch 32768, dim 512, bf16", "n_gpus": world,
                          "ms_per_step": float(ms), "samples_per_s": n_glob / (float(ms) * 1e-3),
                          "alg_tflops_per_gpu": 6.0 * (n_glob / world) * n_glob * dim / (float(ms) * 1e-3) / 1e12,
                          "loss_rank0": float(loss.detach())}))
    if world > 1:
        dist.destroy_process_group()
