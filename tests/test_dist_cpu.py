"""world_size-2/4 gloo tests (CPU) of the multi-rank HOST logic: which collectives
latteclip_b200.loss issues, the label offset, the cross-term / grad-multiplier rules and the
bank all-reduce.  The CUDA entry points are replaced by tests/_abi_double.py (a torch
restatement of the C-ABI contract) -- the kernels themselves are covered by the -m gpu tests.
Expected values: the real gloo run of the reference ClipLoss (tests/golden/clip_dist_w*.npz)."""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, one_sweep, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import _abi_double
    import latteclip_b200 as lb
    from latteclip_b200 import _lib, prototypes as P
    _lib.clip_fwd = _abi_double.clip_fwd
    _lib.clip_bwd = _abi_double.clip_bwd
    _lib.clip_fwd_rows = _abi_double.clip_fwd_rows
    _lib.clip_fwd_cols = _abi_double.clip_fwd_cols
    _lib.rank_sweep_supported = (lambda dtype, dim: bool(one_sweep))
    os.environ["LATTE_B200_BWD_SWEEPS"] = "2" if one_sweep == "fwd_only" else "1"
    _lib.bank_accumulate = _abi_double.bank_accumulate
    _lib.bank_finalize = _abi_double.bank_finalize
    g = np.load(os.path.join(HERE, "golden", f"clip_dist_w{world}.npz"))
    i_all, t_all = torch.from_numpy(g["I"]), torch.from_numpy(g["T"])
    n = i_all.shape[0] // world
    out = {}
    for local_loss in (False, True):
        for gwg in (False, True):
            il = i_all[rank * n:(rank + 1) * n].clone().requires_grad_(True)
            tl = t_all[rank * n:(rank + 1) * n].clone().requires_grad_(True)
            s = torch.tensor(float(g["scale"]), dtype=torch.float64, requires_grad=True)
            mod = lb.ClipLoss(local_loss=local_loss, gather_with_grad=gwg, cache_labels=True,
                              rank=rank, world_size=world)
            # forward() insists on CUDA tensors only through _lib; the double accepts CPU ones
            loss = lb.loss._FusedClipLoss.apply(il, tl, s, local_loss, gwg, rank, world, None, il.dtype, False)
            loss.backward()
            key = f"ll{int(local_loss)}_gwg{int(gwg)}"
            out[key] = dict(loss=float(loss), dI=il.grad.numpy(), dT=tl.grad.numpy(), ds=float(s.grad))
            # gather_features keeps the reference's contract
            ai, at = lb.gather_features(il.detach().requires_grad_(True), tl.detach(), local_loss, gwg, rank, world)
            assert torch.equal(ai.detach(), i_all) and torch.equal(at.detach(), t_all)
            assert ai.requires_grad == (gwg or not local_loss)
    # bank update: every rank must end with the single-process result on the concatenated batch
    gen = torch.Generator().manual_seed(5)
    c, d, b = 6, 16, 8 * world
    bank0 = torch.nn.functional.normalize(torch.randn(c, d, generator=gen), dim=1)
    preds = torch.randint(0, c - 1, (b,), generator=gen)      # class c-1 stays untouched
    zs = torch.randint(0, c - 1, (b,), generator=gen)
    t_ft, t_zs = torch.randn(b, d, generator=gen), torch.randn(b, d, generator=gen)
    sl = slice(rank * 8, (rank + 1) * 8)
    bank = bank0.clone()
    P.update_bank(bank, preds[sl], zs[sl], t_ft[sl], t_zs[sl], world_size=world)
    single = _abi_double.bank_finalize(*_abi_double.bank_accumulate(t_ft, t_zs, preds, zs, c), bank0.clone())
    out["bank_err"] = float((bank - single).abs().max())
    out["bank_untouched"] = bool(torch.equal(bank[c - 1], bank0[c - 1]))
    ret[rank] = out
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,port,one_sweep", [(2, 29721, False), (4, 29723, False),
                                                   (2, 29725, True), (4, 29727, True),
                                                   (2, 29729, "fwd_only"), (4, 29731, "fwd_only")])
def test_multirank_host_logic_matches_gloo_reference(world, port, one_sweep):
    """one_sweep=False: every rank sweeps its row and its column block, LSE vectors exchanged in
    backward.  one_sweep=True: one sweep per rank, column partials all-gathered in forward and
    the text gradient reduce-scattered in backward (the path 16-bit features with dim <= 512
    take on the GPU).  one_sweep="fwd_only": that forward, but the backward sweeps rows and columns
    and exchanges nothing (the default from 8 ranks on)."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, one_sweep, ret), nprocs=world, join=True)
    g = np.load(os.path.join(HERE, "golden", f"clip_dist_w{world}.npz"))
    for r in range(world):
        for key in ("ll0_gwg0", "ll0_gwg1", "ll1_gwg0", "ll1_gwg1"):
            got = ret[r][key]
            assert abs(got["loss"] - float(g[f"{key}_r{r}_loss"])) < 1e-5, (key, r)
            for nm in ("dI", "dT"):
                ref = g[f"{key}_r{r}_{nm}"]
                err = np.linalg.norm(got[nm] - ref) / max(np.linalg.norm(ref), 1e-30)
                assert err < 2e-5, (key, r, nm, err)   # LSE vectors cross the ABI as fp32
            ref_ds = float(g[f"{key}_r{r}_ds"])
            if one_sweep is True and key == "ll1_gwg1":
                # rows-of-this-rank x all-columns partition of the same global sum: only the
                # sum over ranks (what DDP's all-reduce of the parameter gradient sees) matches
                got_sum = sum(ret[q][key]["ds"] for q in range(world))
                ref_sum = sum(float(g[f"{key}_r{q}_ds"]) for q in range(world))
                assert abs(got_sum - ref_sum) < 2e-5 * max(1.0, abs(ref_sum)), (key, r)
            else:
                assert abs(got["ds"] - ref_ds) < 2e-5 * max(1.0, abs(ref_ds)), (key, r)
        assert ret[r]["bank_err"] < 1e-6 and ret[r]["bank_untouched"]


def _siglip_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import _abi_double
    import latteclip_b200 as lb
    from latteclip_b200 import _lib
    _lib.siglip_fwd = _abi_double.siglip_fwd
    _lib.siglip_bwd = _abi_double.siglip_bwd
    g = np.load(os.path.join(HERE, "golden", "siglip.npz"))
    i_all, t_all = torch.from_numpy(g[f"w{world}_I"]), torch.from_numpy(g[f"w{world}_T"])
    n = i_all.shape[0] // world
    il = i_all[rank * n:(rank + 1) * n].clone().requires_grad_(True)
    tl = t_all[rank * n:(rank + 1) * n].clone().requires_grad_(True)
    s = torch.tensor(float(g[f"w{world}_scale"]), dtype=torch.float64, requires_grad=True)
    b = torch.tensor(float(g[f"w{world}_bias"]), dtype=torch.float64, requires_grad=True)
    loss = lb.siglip._FusedSigLip.apply(il, tl, s, b, rank, world, None)
    loss.backward()
    ret[rank] = dict(loss=float(loss), dI=il.grad.numpy(), dT=tl.grad.numpy(), ds=float(s.grad),
                     db=float(b.grad))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,port", [(2, 29741), (3, 29743), (4, 29745)])
def test_siglip_host_logic_matches_gloo_ring_reference(world, port):
    """One text all-gather + reduce-scatter of the text-side product against the reference's
    neighbour-exchange ring (loss.py:521-558) recorded on gloo."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_siglip_worker, args=(world, port, ret), nprocs=world, join=True)
    g = np.load(os.path.join(HERE, "golden", "siglip.npz"))
    for r in range(world):
        got = ret[r]
        ref_loss = float(g[f"w{world}_r{r}_loss"])
        assert abs(got["loss"] - ref_loss) < 1e-5 * abs(ref_loss)            # fp32 across the ABI
        for nm in ("dI", "dT"):
            ref = g[f"w{world}_r{r}_{nm}"]
            err = np.linalg.norm(got[nm] - ref) / np.linalg.norm(ref)
            assert err < 2e-5, (r, nm, err)
        for nm in ("ds", "db"):
            ref = float(g[f"w{world}_r{r}_{nm}"])
            assert abs(got[nm] - ref) < 2e-5 * max(1.0, abs(ref)), (r, nm)
