"""Multi-GPU (NCCL) parity test of the fused ClipLoss + bank update: one process per GPU.
Skipped when fewer than 2 GPUs are visible (the host logic is covered on CPU by
test_dist_cpu.py with gloo)."""

import math
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _sigma(d):
    """Pair noise that leaves the loss at a few tenths (a nearly converged batch would make the
    relative loss / d logit_scale checks vacuous)."""
    return 5.0 if d <= 512 else 6.0


def _worlds():
    """Every world size of {2, 4, 8} that fits the visible GPUs (LATTE_TEST_WORLDS=8 or 2,8
    restricts the list, e.g. to keep an 8-GPU lease short)."""
    fit = [w for w in (2, 4, 8) if w <= torch.cuda.device_count()]
    env = os.environ.get("LATTE_TEST_WORLDS")
    if env:
        want = {int(x) for x in env.split(",") if x.strip()}
        fit = [w for w in fit if w in want]
    return fit


def _worker(rank, world, port, n, d, dtype_name, p2p, ret):
    sys.path.insert(0, ROOT)
    os.environ["LATTE_B200_NO_P2P"] = "0" if p2p else "1"
    os.environ["LATTE_B200_BWD_SWEEPS"] = "2" if p2p == "two_bwd_sweeps" else "1"
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import latteclip_b200 as lb
    from latteclip_b200 import prototypes as P
    dtype = getattr(torch, dtype_name)
    g = torch.Generator().manual_seed(123)
    i_all = F.normalize(torch.randn(n * world, d, generator=g), dim=1)
    sigma = _sigma(d)
    t_all = F.normalize(i_all + sigma * torch.randn(n * world, d, generator=g) / math.sqrt(d), dim=1)
    out = {}
    for local_loss, gwg in ((True, True), (True, False), (False, True), (False, False)):
        il = i_all[rank * n:(rank + 1) * n].to(dev).to(dtype).requires_grad_(True)
        tl = t_all[rank * n:(rank + 1) * n].to(dev).to(dtype).requires_grad_(True)
        s = torch.tensor(100.0, device=dev, requires_grad=True)
        loss = lb.ClipLoss(local_loss=local_loss, gather_with_grad=gwg, rank=rank, world_size=world)(il, tl, s)
        loss.backward()
        out[(local_loss, gwg)] = dict(loss=float(loss.detach()), dI=il.grad.float().cpu(),
                                      dT=tl.grad.float().cpu(), ds=float(s.grad))
    if dtype != torch.float32:
        # SigLipLoss: text all-gather, one sweep, text gradient reduce-scattered (P2P or NCCL)
        il = i_all[rank * n:(rank + 1) * n].to(dev).to(dtype).requires_grad_(True)
        tl = t_all[rank * n:(rank + 1) * n].to(dev).to(dtype).requires_grad_(True)
        s = torch.tensor(20.0, device=dev, requires_grad=True)
        b = torch.tensor(-6.0, device=dev, requires_grad=True)
        loss = lb.SigLipLoss(rank=rank, world_size=world)(il, tl, s, b)
        loss.backward()
        out["siglip"] = dict(loss=float(loss.detach()), dI=il.grad.float().cpu(), dT=tl.grad.float().cpu(),
                             ds=float(s.grad), db=float(b.grad))
    c = 11
    bank = F.normalize(torch.randn(c, d, generator=g), dim=1)
    preds = torch.randint(0, c, (n * world,), generator=g)
    zs = torch.randint(0, c, (n * world,), generator=g)
    sl = slice(rank * n, (rank + 1) * n)
    bank_g = bank.to(dev).clone()
    P.update_bank(bank_g, preds[sl].to(dev), zs[sl].to(dev), t_all[sl].to(dev), i_all[sl].to(dev), world_size=world)
    out["bank"] = bank_g.cpu()
    ret[rank] = out
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("dtype_name,n,d,p2p", [("bfloat16", 512, 512, True), ("bfloat16", 512, 512, False),
                                                 ("bfloat16", 300, 768, True), ("float32", 96, 64, True),
                                                 ("bfloat16", 512, 512, "two_bwd_sweeps")])
def test_nccl_cliploss_matches_oracle(dtype_name, n, d, p2p, world):
    """p2p=True: feature gather and text-gradient reduce-scatter through peer-mapped buffers
    (latte_push_shards, fused GEMM + reduce-scatter); p2p=False: the same one-sweep flow over
    NCCL all-gather / reduce-scatter; "two_bwd_sweeps": P2P gather and one forward sweep, but the
    backward sweeps rows and columns and exchanges nothing (the default from 8 ranks on).
    fp32 features take the two-sweep flow throughout."""
    if world not in _worlds():
        pytest.skip(f"needs {world} GPUs (visible: {torch.cuda.device_count()}, "
                    f"LATTE_TEST_WORLDS={os.environ.get('LATTE_TEST_WORLDS', '')})")
    import oracle
    from oracle.clip_loss import clip_loss_all_ranks
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29801 + n % 7 + 10 * int(p2p is True) + 20 * int(p2p == 'two_bwd_sweeps') + d % 5 + 40 * world, n, d, dtype_name, p2p, ret),
             nprocs=world, join=True)
    dtype = getattr(torch, dtype_name)
    g = torch.Generator().manual_seed(123)
    i_all = F.normalize(torch.randn(n * world, d, generator=g), dim=1)
    sigma = _sigma(d)
    t_all = F.normalize(i_all + sigma * torch.randn(n * world, d, generator=g) / math.sqrt(d), dim=1)
    ir, tr = i_all.to(dtype).float(), t_all.to(dtype).float()
    ish = [ir[r * n:(r + 1) * n] for r in range(world)]
    tsh = [tr[r * n:(r + 1) * n] for r in range(world)]
    gtol = 2.6e-3 if dtype_name == "bfloat16" else 5e-5   # bf16 outputs: see test_gpu_clip.py
    report = []
    for key in ((True, True), (True, False), (False, True), (False, False)):
        lo, di, dt, ds = clip_loss_all_ranks(ish, tsh, 100.0, key[0], key[1])
        for r in range(world):
            got = ret[r][key]
            e_i = float((got["dI"].double() - di[r]).norm() / di[r].norm())
            e_t = float((got["dT"].double() - dt[r]).norm() / dt[r].norm())
            report.append((key, r, got["loss"], float(lo[r]), e_i, e_t, float(di[r].norm()), got["ds"], float(ds[r])))
    for row in report:
        print("local_loss=%s gwg=%s rank %d loss %.6f ref %.6f dI %.3e dT %.3e |dI| %.3e ds %.4e ref %.4e" %
              (row[0][0], row[0][1], *row[1:]))
    one_sweep = dtype_name == "bfloat16"     # 16-bit, dim <= 512: one logit sweep per rank
    for key, r, loss, lref, e_i, e_t, _, dsv, dsr in report:
        assert lref > 0.05, "test data must stay away from convergence"
        assert abs(loss - lref) <= 5e-5 * abs(lref), (key, r, loss, lref)
        assert e_i < gtol and e_t < gtol, (key, r, e_i, e_t)
        if one_sweep and key == (True, True):
            # d loss / d s covers rows-of-this-rank x all columns: the sum over ranks (what
            # DDP's all-reduce of the parameter gradient sees) is the reference's
            got = sum(x[7] for x in report if x[0] == key)
            ref_sum = sum(x[8] for x in report if x[0] == key)
            assert abs(got - ref_sum) <= 2e-3 * abs(ref_sum), (key, r, got, ref_sum)
        else:
            assert abs(dsv - dsr) <= 2e-3 * abs(dsr), (key, r, dsv, dsr)
    if dtype_name != "float32":
        from oracle.siglip import siglip_all_ranks
        lo, di, dt, ds, db = siglip_all_ranks(ish, tsh, 20.0, -6.0)
        for r in range(world):
            got = ret[r]["siglip"]
            assert abs(got["loss"] - float(lo[r])) <= 1e-5 * abs(float(lo[r])), ("siglip", r)
            e_i = float((got["dI"].double() - di[r]).norm() / di[r].norm())
            e_t = float((got["dT"].double() - dt[r]).norm() / dt[r].norm())
            print("siglip rank %d loss %.6f ref %.6f dI %.3e dT %.3e ds %.4e ref %.4e db %.4e ref %.4e" %
                  (r, got["loss"], float(lo[r]), e_i, e_t, got["ds"], float(ds[r]), got["db"], float(db[r])))
            assert e_i < gtol and e_t < gtol, ("siglip", r, e_i, e_t)
            assert abs(got["ds"] - float(ds[r])) <= 2e-3 * abs(float(ds[r])) + 1e-6
            assert abs(got["db"] - float(db[r])) <= 2e-3 * abs(float(db[r])) + 1e-6
    c = 11
    bank = F.normalize(torch.randn(c, d, generator=g), dim=1)
    preds = torch.randint(0, c, (n * world,), generator=g)
    zs = torch.randint(0, c, (n * world,), generator=g)
    ref = oracle.update_bank(bank, preds, zs, t_all, i_all)
    for r in range(world):
        assert float((ret[r]["bank"] - ref).norm() / ref.norm()) < 1e-5
        assert torch.equal(ret[r]["bank"], ret[0]["bank"])
