#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native LatteCLIP loss head.

Metric (BASELINE.json): ClipLoss forward + backward samples/s at global batch 32768,
dim 512 (bf16 features, logit_scale 100), on 1/2/4/8 B200.  A "step" is one fused ClipLoss
forward + backward over the whole global batch (the reference's loss.py:120-130 + autograd),
sharded by rows over the ranks exactly like open_clip's local_loss / gather_with_grad mode.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
    python bench.py --impl reference [--gpus N] [--steps K] ...     # reference CPU arm

For N > 1 launch with torchrun (one rank per GPU, NCCL).  Rank 0 prints ONE JSON line.
"""

from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch
import torch.nn.functional as F

N_GLOBAL = 32768
DIM = 512
SCALE = 100.0
METRIC = "clip_loss_fwd_bwd_samples_per_sec"
UNIT = "samples/s"
# One CPU row block for both the cpu_baseline leg and the --impl reference arm
CPU_BLOCK_ROWS = 4096
# weak-scaling operating point: rows per GPU
WEAK_ROWS_PER_GPU = 4096


def ncu_traffic_bytes():
    """DRAM traffic of the dominant kernel (pair_gemm_kernel) per launch at N = 1, read from the
    committed ncu capture of the shipped kernel (profiles/gemm_traffic.json: dram__bytes_read.sum +
    dram__bytes_write.sum of one `ncu --set full` launch); None when no capture is committed."""
    path = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    try:
        return float(json.load(open(path))["dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        return None


def synth_shard(n_global, dim, rank, world, set_id=0):
    """Seeded synthetic unit-norm features with correlated pairs:
    I = normalize(randn), T = normalize(I + sigma * randn / sqrt(D)), sigma = 4."""
    n = n_global // world
    g = torch.Generator().manual_seed(1234 + 1000 * 2 + rank + 97 * set_id)
    i = F.normalize(torch.randn(n, dim, generator=g), dim=1)
    t = F.normalize(i + 4.0 * torch.randn(n, dim, generator=g) / math.sqrt(dim), dim=1)
    return i, t


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(burst=float(p["bf16_tflops"]), sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    hbm=float(p["hbm_gbs"]), source="measured (MEASURED_PEAKS.json)")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.

    Two fields per sample -- the SM clock and the event-reason bitmask.  Every NVML read made while the GPU
    is busy stalls it on the boxes of this pool (a 20-step region measured 3.58 ms/step while power.draw +
    clocks + four reason fields were sampled every 50 ms, 3.39 ms/step with five fields every 100 ms, and
    3.16 ms/step in the identical region that followed without the sampler), so the sampler asks for as
    little as the clocks record needs, every 50 ms; clocks.max.sm is static and read once before the load."""
    Q = "timestamp,clocks.sm,clocks_event_reasons.active"
    REASON_BITS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
                   0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.sm_max = None

    def start(self):
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits",
                                  "-i", str(self.gpu)], capture_output=True, text=True, timeout=20).stdout
            self.sm_max = float(out.strip().splitlines()[0])
        except Exception:
            self.sm_max = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def load_start(self):
        """Start of the continuous load (first warm-up step)."""
        self.t_load = time.time()

    def mark(self):
        """Start of the timed region: only samples taken inside it are reported (falls back to the
        samples of the continuous load before it if the region held fewer than two)."""
        self.t_mark = time.time()

    def stop(self):
        t_end = time.time() + 0.02          # a sample is printed a little after it was taken
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        t_mark = getattr(self, "t_mark", 0.0)
        t_load = getattr(self, "t_load", t_mark)
        inside = [ln for (ts, ln) in self.lines if t_mark <= ts <= t_end]
        window = "timed region + the identical repeat region that follows it"
        if len(inside) < 2:
            inside = [ln for (ts, ln) in self.lines if t_load + 0.02 <= ts <= t_end]
            window = "warm-up + timed region + repeat region (continuous load)"
        if not inside:
            inside = [ln for (_, ln) in self.lines[1:]] or [ln for (_, ln) in self.lines]
            window = "whole run"
        self.window = window
        sm, reasons = [], set()
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 3:
                continue
            try:
                sm.append(float(f[1]))
                mask = int(f[2], 16)
            except ValueError:
                continue
            for bit, nm in self.REASON_BITS.items():
                if mask & bit:
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.sm_max,
                "samples": len(sm), "window": self.window, "reasons": sorted(reasons)}


def cpu_model_name():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.lower().startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_reference_sample(steps, warmup):
    """The reference algorithm on the host cores (oracle port, fp32, all threads): forward +
    backward of the loss terms owned by the first CPU_BLOCK_ROWS rows of the N = 32768 problem
    (what one rank of a local_loss run computes, loss.py:108-110).  Per-sample cost equals the full
    batch's, so samples/s = rows / time.  The same block size serves the cpu_baseline leg of our
    arm and the --impl reference arm."""
    from oracle.clip_loss import clip_loss_row_block_sample
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    i, t = synth_shard(N_GLOBAL, DIM, 0, 1)
    rows = CPU_BLOCK_ROWS
    for _ in range(max(warmup, 1)):
        clip_loss_row_block_sample(i, t, SCALE, rows)
    t0 = time.perf_counter()
    for _ in range(steps):
        clip_loss_row_block_sample(i, t, SCALE, rows)
    dt = (time.perf_counter() - t0) / steps
    return dict(value=rows / dt, unit=UNIT, cores=torch.get_num_threads(), cpu_model=cpu_model_name(), kind="port",
                sample=f"fwd+bwd of the first {rows} rows (both CE directions) of the N={N_GLOBAL}, "
                       f"D={DIM} fp32 problem, {steps} steps, {dt * 1e3:.1f} ms/step"), dt


def load_reference_loss_module():
    """The reference's own open_clip/loss.py, unmodified, from baseline/_ref (installed by
    baseline/install_ref.sh with pip --target; the file imports with torch alone).  None when the
    install is absent."""
    path = os.path.join(ROOT, "baseline", "_ref", "open_clip", "loss.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_open_clip_loss", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cpu_reference_module(n=8192, reps=2):
    """The reference's ClipLoss MODULE itself (baseline/_ref, world_size 1) on the host cores at a
    batch it finishes in about a second.  Extra information: a batch of n costs n/N of the headline
    batch per sample, so this is NOT the headline workload (the row-block sample is)."""
    mod = load_reference_loss_module()
    if mod is None:
        return None
    torch.set_num_threads(os.cpu_count() or 1)
    i, t = synth_shard(n, DIM, 0, 1)
    il, tl = i.clone().requires_grad_(True), t.clone().requires_grad_(True)
    s = torch.tensor(SCALE, requires_grad=True)
    loss_fn = mod.ClipLoss()
    loss_fn(il, tl, s).backward()
    t0 = time.perf_counter()
    for _ in range(reps):
        il.grad = tl.grad = s.grad = None
        loss = loss_fn(il, tl, s)
        loss.backward()
    dt = (time.perf_counter() - t0) / reps
    return {"batch": n, "ms_per_step": dt * 1e3, "samples_per_s": n / dt, "loss": float(loss.detach()),
            "kind": "reference (baseline/_ref/open_clip/loss.py ClipLoss, unmodified)"}


def prototype_kernel_rates(dev, peaks, batch=N_GLOBAL, dim=DIM, classes=47):
    """Achieved HBM GB/s of the prototype-path kernels at the headline batch (fp32 features,
    47 DTD classes), CUDA events over rotating inputs larger than L2.  Algorithmic bytes per
    launch as in SURVEY 8d."""
    from latteclip_b200 import _lib
    g = torch.Generator().manual_seed(4321)
    sets = 6
    bank = F.normalize(torch.randn(classes, dim, generator=g), dim=1).to(dev)
    cls = F.normalize(torch.randn(classes, dim, generator=g), dim=1).to(dev)
    xs = [F.normalize(torch.randn(batch, dim, generator=g), dim=1).to(dev) for _ in range(sets)]
    ps = [F.normalize(torch.randn(batch, dim, generator=g), dim=1).to(dev) for _ in range(sets)]
    preds = torch.randint(0, classes, (batch,), generator=g).to(dev)
    zs = torch.randint(0, classes, (batch,), generator=g).to(dev)
    w = [torch.rand(batch, generator=g).to(dev) + 0.1 for _ in range(4)]

    def timeit(fn, reps=12):
        # `reps` calls captured in ONE CUDA graph and replayed between two events on the replay
        # stream: the kernels are 20-160 us, the host side of a call (ctypes, tensor-map encoding,
        # output allocation) ~20 us, so an eager loop would time the host for the short ones
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for k in range(3):
                fn(k)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for k in range(reps):
                fn(k)
        graph.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    fb = 4
    cp = (classes + 15) // 16 * 16
    planes_bytes = 2 * 3 * cp * dim * 2
    cls_planes = _lib.nxc_split_prototypes(bank, normalize=True)
    snap_planes = _lib.nxc_split_prototypes(bank, normalize=False)
    xh = [x.bfloat16() for x in xs]
    clsh = cls.bfloat16()

    def step_jobs(feats, ct, k):
        # the four products of one LatteCLIP step (prototypes.step_similarities): pseudo-label argmax
        # on the images, top-2 margins of the image-description, group-description and class-name texts
        return [dict(x=feats[k % sets], planes=cls_planes, scale=100.0, argmax=True),
                dict(x=feats[(k + 1) % sets], planes=snap_planes, margin=True),
                dict(x=feats[(k + 2) % sets], planes=snap_planes, margin=True),
                dict(x=ct, planes=snap_planes, margin=True)]
    rows = [
        ("nxc_stream_kernel<convert>, fp32 features: the step's 4 stacked N x C products in one launch "
         "(train.py:410-411 argmax + 3 compute_text_weights margins :292-303)",
         (3 * batch + classes) * dim * 4 + planes_bytes + batch * 16,
         lambda k: _lib.nxc_multi(step_jobs(xs, cls, k))),
        ("nxc_stream_kernel<direct>, bf16 features: the same 4 stacked products",
         (3 * batch + classes) * dim * 2 + planes_bytes + batch * 16,
         lambda k: _lib.nxc_multi(step_jobs(xh, clsh, k))),
        ("nxc_stream_kernel<convert>, fp32 features: one product (pseudo-label argmax only)",
         batch * dim * 4 + planes_bytes // 2 + batch * 8,
         lambda k: _lib.nxc_multi(step_jobs(xs, cls, k)[:1])),
        ("nxc_tc_kernel, fp32 features: per-product path kept for C > 64 (argmax)",
         batch * dim * fb + classes * dim * 4 + batch * 8,
         lambda k: _lib.nxc_argmax_margin(xs[k % sets], bank, scale=100.0, want_argmax=True, want_margin=False)),
        # per_image + per_group read, t_ft + t_zs written; the label-text / bank rows are gathers
        # from [C, D] tables (96 KB each at C = 47) that stay in L2
        ("mix_ema_fwd_kernel (train.py:472-488)",
         4 * batch * dim * fb + 2 * classes * dim * 4 + 6 * batch * 4 + 2 * batch * 8,
         lambda k: _lib.mix_ema_fwd(cls, xs[k % sets], ps[k % sets], bank, preds, zs, w[0], w[1], w[2], w[3], 0.01, "row")),
        ("mix_ema_bwd (rows kernel + per-class segment sums)",
         4 * batch * dim * fb + classes * dim * 4 + 6 * batch * 4,
         lambda k: _lib.mix_ema_bwd(xs[k % sets], ps[k % sets], preds, zs, w[0], w[1], w[2], w[3], 0.01, "row", classes)),
        ("bank_accumulate (per-class segment sums, train.py:508-526)",
         2 * batch * dim * fb + 2 * batch * 8 + classes * dim * 4,
         lambda k: _lib.bank_accumulate(xs[k % sets], ps[k % sets], preds, zs, classes)),
    ]
    out = []
    for name, nbytes, fn in rows:
        ms = timeit(fn)
        gbs = nbytes / (ms * 1e-3) / 1e9
        out.append({"kernel": name, "ms": ms, "alg_bytes": nbytes, "achieved_gbs": gbs,
                    "frac_of_measured_hbm": gbs / peaks["hbm"]})
    return {"batch": batch, "dim": dim, "classes": classes, "dtype": "f32 unless the row says bf16",
            "timing": "12 calls captured in one CUDA graph, replay timed with events; 6 rotating input "
                      "sets (403 MiB fp32) larger than the L2", "rows": out}


def siglip_times(dev, n=None, dim=None, reps=10):
    """SigLipLoss fwd + bwd (open_clip loss.py:453-560) at the headline shape on one GPU through the
    drop-in module, split into forward sweep and backward (sweep + gradient GEMMs) by CUDA events.
    Extra information beside the headline metric (same 6 n N D algorithmic FLOP per call)."""
    import latteclip_b200 as lb
    n = n or N_GLOBAL
    dim = dim or DIM
    i, t = synth_shard(n, dim, 0, 1, set_id=9)
    il = i.to(dev).bfloat16().requires_grad_(True)
    tl = t.to(dev).bfloat16().requires_grad_(True)
    s = torch.tensor(10.0, device=dev, requires_grad=True)      # SigLIP init operating point,
    b = torch.tensor(-10.0, device=dev, requires_grad=True)     # training/main.py:225-227
    mod = lb.SigLipLoss()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    fwd_ms = bwd_ms = 0.0
    last = 0.0
    for k in range(3 + reps):
        for x in (il, tl, s, b):
            x.grad = None
        ev[0].record()
        loss = mod(il, tl, s, b)
        ev[1].record()
        loss.backward()
        ev[2].record()
        torch.cuda.synchronize()
        if k >= 3:
            fwd_ms += ev[0].elapsed_time(ev[1]) / reps
            bwd_ms += ev[1].elapsed_time(ev[2]) / reps
        last = float(loss.detach())
    ms = fwd_ms + bwd_ms
    return {"workload": f"SigLipLoss fwd+bwd, batch {n}, dim {dim}, bf16 features, scale 10, bias -10",
            "ms_per_step": ms, "fwd_ms": fwd_ms, "bwd_ms": bwd_ms, "samples_per_s": n / (ms * 1e-3),
            "alg_tflops": 6.0 * n * n * dim / (ms * 1e-3) / 1e12, "loss": last}


def latteclip_head_times(dev, batch=512, dim=512, classes=47, reps=20, axis="quirk", cpu=True):
    """One LatteCLIP head step (train.py:384-530: pseudo-labels, margins, mixture + EMA, two
    ClipLoss calls sharing the image features, backward, bank update) at the reference's own
    fine-tuning shape (BASELINE cfg2: batch 512, dim 512, 47 classes, quirk label broadcast), on
    the GPU through latteclip_b200.prototypes, and the oracle port of the same step on the host
    cores.  Extra information beside the headline metric."""
    import latteclip_b200 as lb
    from latteclip_b200 import prototypes as P
    g = torch.Generator().manual_seed(77)
    bank = F.normalize(torch.randn(classes, dim, generator=g), dim=1)
    cls = F.normalize(bank + 0.3 * torch.randn(classes, dim, generator=g), dim=1)
    true = torch.randint(0, classes, (batch,), generator=g)
    mk = lambda s: F.normalize(bank[true] + s * torch.randn(batch, dim, generator=g) * 3 / dim ** 0.5, dim=1)
    img, pimg, pgrp = mk(1.2), mk(0.9), mk(0.7)
    zs = torch.randint(0, classes, (batch,), generator=g)
    loss_fn = lb.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True)
    dimg, dcls, dpi, dpg = (x.to(dev).bfloat16().requires_grad_(True) for x in (img, cls, pimg, pgrp))
    log_s = torch.tensor(math.log(SCALE), device=dev, requires_grad=True)
    bank_d, snap_d, zs_d = bank.to(dev).clone(), bank.to(dev).clone(), zs.to(dev)

    def step():
        for x in (dimg, dcls, dpi, dpg, log_s):
            x.grad = None
        out = P.prototype_step(dimg, log_s.exp(), bank_d, snap_d, zs_d, dcls, dpi, dpg, loss_fn,
                               alpha=0.01, label_weight_axis=axis)
        out["loss"].backward()
        P.update_bank(bank_d, out["preds"], zs_d, out["t_ft"], out["t_zs"])

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        step()
    b.record()
    torch.cuda.synchronize()
    eager_ms = a.elapsed_time(b) / reps
    # the same step through prototypes.GraphedPrototypeStep: captured once in a CUDA graph, replayed per
    # step; the inputs are copied into the graph's static buffers and the gradients handed back to
    # autograd every step (that is inside the timed region)
    graphed = P.GraphedPrototypeStep(loss_fn, alpha=0.01, label_weight_axis=axis)

    def gstep():
        for x in (dimg, dcls, dpi, dpg, log_s):
            x.grad = None
        out = graphed(dimg, log_s.exp(), bank_d, snap_d, zs_d, dcls, dpi, dpg)
        out["loss"].backward()

    for _ in range(3):
        gstep()
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        gstep()
    b.record()
    torch.cuda.synchronize()
    ours_ms = a.elapsed_time(b) / reps
    res = {"workload": f"LatteCLIP head step, batch {batch}, dim {dim}, {classes} classes, bf16 features, "
                       f"label_weight_axis={axis}",
           "ms_per_step": ours_ms, "samples_per_s": batch / (ours_ms * 1e-3),
           "path": "prototypes.GraphedPrototypeStep (CUDA-graph replay of prototype_step + backward + update_bank)",
           "eager_launch_ms_per_step": eager_ms}
    if not cpu:
        return res
    import oracle
    torch.set_num_threads(os.cpu_count() or 1)
    ci, cc, cpi, cpg = (x.clone().requires_grad_(True) for x in (img, cls, pimg, pgrp))
    cl = torch.tensor(math.log(SCALE), requires_grad=True)

    def cpu_step():
        out = oracle.prototype_step(ci, cl.exp(), bank.clone(), bank.clone(), zs, cc, cpi, cpg, alpha=0.01,
                                    label_weight_axis=axis)
        out["loss"].backward()

    cpu_step()
    t0 = time.perf_counter()
    for _ in range(3):
        cpu_step()
    cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
    res.update(cpu_oracle_ms_per_step=cpu_ms, cpu_cores=torch.get_num_threads())
    return res


def gpu_eager_reference(dev, reps=3):
    """The reference's GPU path restated op for op in eager PyTorch on this GPU (loss.py:102-130:
    two logit GEMMs with the scale on the A operand, two F.cross_entropy over materialised [N, N]
    logits, autograd backward) at the headline shape -- the like-for-like bar SURVEY 8d asks for,
    reported beside the CPU measurement of the reference arm.  None of our kernels run here."""
    i, t = synth_shard(N_GLOBAL, DIM, 0, 1)
    ref_mod = load_reference_loss_module()
    out = {"implementation": "baseline/_ref/open_clip/loss.py ClipLoss (unmodified reference module)"
           if ref_mod is not None else "restatement of loss.py:102-130 (baseline/_ref not installed)"}
    for name, ac in (("amp_bf16", torch.bfloat16), ("fp32", None)):
        try:
            il = i.to(dev).requires_grad_(True)
            tl = t.to(dev).requires_grad_(True)
            s = torch.tensor(SCALE, device=dev, requires_grad=True)
            labels = torch.arange(N_GLOBAL, device=dev)
            ref_loss = ref_mod.ClipLoss(cache_labels=True).to(dev) if ref_mod is not None else None

            def step():
                il.grad = tl.grad = s.grad = None
                with torch.autocast("cuda", dtype=ac or torch.bfloat16, enabled=ac is not None):
                    if ref_loss is not None:
                        loss = ref_loss(il, tl, s)
                    else:
                        logits_per_image = s * il @ tl.T
                        logits_per_text = s * tl @ il.T
                        loss = (F.cross_entropy(logits_per_image, labels) + F.cross_entropy(logits_per_text, labels)) / 2
                loss.backward()
                return loss

            torch.cuda.reset_peak_memory_stats(dev)
            step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                loss = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            out[name] = {"ms_per_step": ms, "samples_per_s": N_GLOBAL / (ms * 1e-3), "loss": float(loss.detach()),
                         "peak_mem_gib": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
            del il, tl, s, loss
            torch.cuda.empty_cache()
        except Exception as exc:      # e.g. out of memory on a smaller part: report, do not fail the arm
            out[name] = {"error": str(exc).splitlines()[0][:200]}
    return out


def gpu_eager_head_reference(dev, batch=512, dim=512, classes=47, reps=10):
    """The loss-head part of one ``train_one_epoch_v2`` iteration restated op for op in eager PyTorch on
    this GPU, at the reference's own fine-tuning shape (BASELINE cfg2), INCLUDING its per-sample Python
    loops and their implicit device-to-host reads (train.py:412-431 label gathers, :508-530 bank update),
    six ``compute_text_weights`` calls (:444-449, bmm + topk :292-303), the mixture (:472-488, literal
    broadcast: B == D) and two calls of the reference's own ClipLoss module (baseline/_ref) sharing the
    image features (:491-504), then backward (:506).  Towers are outside: features arrive as tensors;
    the class-name text features are gathered from a per-class table instead of re-encoded.  None of our
    kernels run here; it is the like-for-like bar of `latteclip_head` in our arm."""
    import collections
    ref_mod = load_reference_loss_module()
    g = torch.Generator().manual_seed(77)
    bank = F.normalize(torch.randn(classes, dim, generator=g), dim=1)
    cls = F.normalize(bank + 0.3 * torch.randn(classes, dim, generator=g), dim=1)
    true = torch.randint(0, classes, (batch,), generator=g)
    mk = lambda sd: F.normalize(bank[true] + sd * torch.randn(batch, dim, generator=g) * 3 / dim ** 0.5, dim=1)  # noqa: E731
    img, pimg, pgrp = mk(1.2), mk(0.9), mk(0.7)
    zs_ids = torch.randint(0, classes, (batch,), generator=g).tolist()
    names = [f"class{k}" for k in range(classes)]
    name2id = {c: k for k, c in enumerate(names)}
    zs_names = [(names[k],) for k in zs_ids]
    memory_bank = {c: bank[k].to(dev).clone() for k, c in enumerate(names)}
    dimg, dcls, dpi, dpg = (x.to(dev).requires_grad_(True) for x in (img, cls, pimg, pgrp))
    log_s = torch.tensor(math.log(SCALE), device=dev, requires_grad=True)
    prototypes = torch.stack([memory_bank[c] for c in names])                    # train.py:347-350
    loss_mod = ref_mod.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True) if ref_mod else None
    labels = torch.arange(batch, device=dev)

    def clip(i, t, s):
        if loss_mod is not None:
            return loss_mod(i, t, s)
        return (F.cross_entropy(s * i @ t.T, labels) + F.cross_entropy(s * t @ i.T, labels)) / 2

    def text_weights(x, protos, preds):                                          # train.py:292-303
        w = torch.bmm(x.unsqueeze(1), protos.T.unsqueeze(0).expand(x.shape[0], -1, -1)).squeeze(1)
        top2, idx = torch.topk(w, 2, dim=1)
        _ = idx[:, 0] == preds
        return top2[:, 0] - top2[:, 1]

    def step():
        for x in (dimg, dcls, dpi, dpg, log_s):
            x.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            classifier = F.normalize(torch.stack([memory_bank[c] for c in names]), dim=1).T      # :384-389
            logit_scale = log_s.exp()
            preds = (100.0 * dimg @ classifier).argmax(dim=1)                                     # :410-411
            zs_preds = torch.zeros_like(preds)
            mb, mbz, lab, labz = [], [], [], []
            for i in range(batch):                                                                # :415-431
                zname = zs_names[i][0]
                zs_preds[i] = name2id[zname]
                cname = names[preds[i]]                     # implicit device-to-host read per sample
                lab.append(name2id[cname]); labz.append(name2id[zname])
                mb.append(memory_bank[cname]); mbz.append(memory_bank[zname])
            mb, mbz = torch.stack(mb), torch.stack(mbz)
            l_ft, l_zs = dcls[torch.tensor(lab, device=dev)], dcls[torch.tensor(labz, device=dev)]
            w_img = text_weights(dpi, prototypes, preds).detach() + 1e-6                          # :444-449
            w_grp = text_weights(dpg, prototypes, preds).detach() + 1e-6
            w_img_z = text_weights(dpi, prototypes, zs_preds).detach() + 1e-6
            w_grp_z = text_weights(dpg, prototypes, zs_preds).detach() + 1e-6
            w_lbl = text_weights(l_ft, prototypes, preds).detach() + 1e-6
            w_lbl_z = text_weights(l_zs, prototypes, zs_preds).detach() + 1e-6
            tot, tot_z = w_lbl + w_img + w_grp, w_lbl_z + w_img_z + w_grp_z                       # :472-473
            t_ft = (w_lbl * l_ft + dpi * w_img.unsqueeze(1) + dpg * w_grp.unsqueeze(1)) / tot.unsqueeze(1)
            t_zs = (w_lbl * l_zs + dpi * w_img_z.unsqueeze(1) + dpg * w_grp_z.unsqueeze(1)) / tot_z.unsqueeze(1)
            t_ft = mb + 0.01 * (t_ft - mb)                                                        # :487-488
            t_zs = mbz + 0.01 * (t_zs - mbz)
            total = clip(dimg, t_ft, logit_scale) + clip(dimg, t_zs, logit_scale)                 # :491-504
        total.backward()                                                                          # :506
        with torch.no_grad():                                                                     # :508-530
            temp, cnt = {}, collections.defaultdict(int)
            for i in range(batch):
                pname, zname = names[preds[i]], names[zs_preds[i]]
                for nm in (zname, pname):
                    if nm not in temp:
                        temp[nm] = torch.zeros_like(t_ft[i])
                temp[zname] += t_zs[i]
                temp[pname] += t_ft[i]
                cnt[zname] += 1
                cnt[pname] += 1
            for nm in temp:
                memory_bank[nm] = F.normalize(temp[nm] / cnt[nm], dim=0)
        return total

    step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        loss = step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    return {"workload": f"LatteCLIP head step, batch {batch}, dim {dim}, {classes} classes, eager PyTorch with the "
                        "reference's per-sample loops, bf16 autocast",
            "ms_per_step": ms, "samples_per_s": batch / (ms * 1e-3), "loss": float(loss.detach()),
            "clip_loss": "baseline/_ref/open_clip/loss.py ClipLoss" if ref_mod else "restatement of loss.py:102-130"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, dt = cpu_reference_sample(max(args.steps, 1), min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"ClipLoss fwd+bwd, global batch {N_GLOBAL}, dim {DIM}, logit_scale {SCALE:g}",
                   "note": "reference algorithm (oracle port of open_clip/loss.py) on the host CPU cores"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    module = cpu_reference_module()
    if module is not None:
        line["reference_module_cpu"] = module
    if torch.cuda.is_available():
        line["gpu_eager"] = dict(gpu_eager_reference(torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))),
                                 note="the same algorithm in eager PyTorch on this GPU (materialised logits); "
                                      "extra information, the arm's value is the CPU measurement")
        try:
            line["gpu_eager_head"] = gpu_eager_head_reference(
                torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
        except Exception as exc:
            line["gpu_eager_head"] = {"error": str(exc).splitlines()[0][:200]}
    print(json.dumps(line), flush=True)


def fp64_reference(all_i, all_t, scale, n_loc, off, rows_idx, chunk=2048):
    """fp64 torch restatement of loss.py:102-130 + autograd on the GPU for the parity check of
    every bench run (the checker, outside every timed region): from the gathered features (the
    same bf16-rounded values the kernels read) -> per-rank loss of the rank that owns rows
    [off, off + n_loc) under local_loss + gather_with_grad, the gradient rows `rows_idx` (global
    indices) of dI and dT, and the GLOBAL d loss / d logit_scale (sum over ranks)."""
    n = all_i.shape[0]
    I = all_i.double()
    T = all_t.double()
    row_lse = torch.empty(n, dtype=torch.float64, device=I.device)
    col_lse = torch.full((n,), -float("inf"), dtype=torch.float64, device=I.device)
    for c0 in range(0, n, chunk):
        S = scale * (I[c0:c0 + chunk] @ T.T)
        row_lse[c0:c0 + chunk] = torch.logsumexp(S, dim=1)
        col_lse = torch.logaddexp(col_lse, torch.logsumexp(S, dim=0))
    diag = scale * (I * T).sum(dim=1)
    sl = slice(off, off + n_loc)
    loss = ((row_lse[sl] - diag[sl]).mean() + (col_lse[sl] - diag[sl]).mean()) / 2
    # d(sum_r L_r)/dS_ij = (P^row_ij + P^col_ij - 2 delta_ij) / (2 n_loc)
    ds = torch.zeros((), dtype=torch.float64, device=I.device)
    for c0 in range(0, n, chunk):
        S = scale * (I[c0:c0 + chunk] @ T.T)
        G = torch.exp(S - row_lse[c0:c0 + chunk, None]) + torch.exp(S - col_lse[None, :])
        G[torch.arange(G.shape[0]), torch.arange(c0, c0 + G.shape[0])] -= 2.0
        ds += (G * S).sum() / scale
    ds = ds / (2.0 * n_loc)
    Sr = scale * (I[rows_idx] @ T.T)                                   # [R, N]
    Gr = torch.exp(Sr - row_lse[rows_idx, None]) + torch.exp(Sr - col_lse[None, :])
    Gr[torch.arange(len(rows_idx)), rows_idx] -= 2.0
    dI = (scale / (2.0 * n_loc)) * (Gr @ T)
    Sc = scale * (T[rows_idx] @ I.T)                                   # [R, N] = S[:, rows]^T
    Gc = torch.exp(Sc - row_lse[None, :]) + torch.exp(Sc - col_lse[rows_idx, None])
    Gc[torch.arange(len(rows_idx)), rows_idx] -= 2.0
    dT = (scale / (2.0 * n_loc)) * (Gc @ I)
    return loss, dI, dT, ds


def parity_check(loss_fn, il, tl, log_s, rank, world, dev, n_rows=64):
    """One step through the product path, compared with fp64_reference on this rank: loss, n_rows
    sampled rows of dI and dT (bf16 outputs: rel <= 2.6e-3, the north_star's 2e-3 on the fp32 values
    plus the 1.63e-3 rms of rounding any gradient to bf16, see tests/test_gpu_clip.py), and the
    rank-summed d loss / d logit_scale (rel <= 2e-3).  Worst case over ranks is reported."""
    import torch.distributed as dist
    il.grad = None; tl.grad = None; log_s.grad = None
    loss = loss_fn(il, tl, log_s.exp())
    loss.backward()
    n_loc = il.shape[0]
    n_all = n_loc * world
    if world > 1:
        all_i = torch.empty(n_all, il.shape[1], dtype=il.dtype, device=dev)
        all_t = torch.empty(n_all, il.shape[1], dtype=il.dtype, device=dev)
        dist.all_gather_into_tensor(all_i, il.detach())
        dist.all_gather_into_tensor(all_t, tl.detach())
    else:
        all_i, all_t = il.detach(), tl.detach()
    off = rank * n_loc
    g = torch.Generator().manual_seed(99 + rank)
    rows = (torch.randperm(n_loc, generator=g)[:n_rows].sort().values + off).to(dev)
    ref_loss, ref_di, ref_dt, ref_ds = fp64_reference(all_i, all_t, SCALE, n_loc, off, rows)
    rel = lambda a, b: float((a.double() - b).norm() / b.norm())
    e_loss = abs(float(loss.detach()) - float(ref_loss)) / abs(float(ref_loss))
    e_di = rel(il.grad[rows - off], ref_di)
    e_dt = rel(tl.grad[rows - off], ref_dt)
    # log_s.grad = s * d loss / d s of this rank's rows x all columns; the sum over ranks is the
    # reference's global sum (what DDP's all-reduce of the parameter gradient sees)
    ds = log_s.grad.detach().double().clone() / SCALE
    if world > 1:
        dist.all_reduce(ds, op=dist.ReduceOp.SUM)
    e_ds = abs(float(ds) - float(ref_ds)) / abs(float(ref_ds))
    errs = torch.tensor([e_loss, e_di, e_dt, e_ds], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    e_loss, e_di, e_dt, e_ds = (float(x) for x in errs)
    loss_tol = 1e-5 + 2e-5 / abs(float(ref_loss))
    ok = e_loss <= loss_tol and e_di <= 2.6e-3 and e_dt <= 2.6e-3 and e_ds <= 2e-3
    return {"loss_rel": e_loss, "dI_rel": e_di, "dT_rel": e_dt, "dscale_rel": e_ds,
            "rows_checked_per_rank": n_rows, "loss_tol": loss_tol, "grad_tol": 2.6e-3,
            "dscale_tol": 2e-3, "reference": "fp64 torch on the gathered bf16 features (all ranks)",
            "ok": bool(ok)}


def weak_scaling_point(lb, rank, world, dev, steps, rows=WEAK_ROWS_PER_GPU):
    """SURVEY 8d's weak-scaling datum: `rows` rows per GPU (global batch = rows * W), per-GPU credited
    FLOP/s = 6 n N D / t, beside the same GPU running the one-rank problem (n = N = rows) so the
    efficiency refers to the same box.  rows = 4096 is SURVEY 8d's operating point (launch-bound on one
    GPU); rows = 32768 keeps the headline per-GPU batch fixed while the global batch grows to 32768 W."""
    import torch.distributed as dist
    n = rows
    n_glob = n * world
    sets = []
    for sidx in range(4):
        i, t = synth_shard(n_glob, DIM, rank, world, 20 + sidx)
        sets.append((i.bfloat16().to(dev).requires_grad_(True), t.bfloat16().to(dev).requires_grad_(True)))
    log_s = torch.tensor(math.log(SCALE), device=dev, requires_grad=True)

    def run(fn, k_steps):
        for k in range(8):                 # every rotating set twice: workspaces, gradient buffers, clocks
            i, t = sets[k % 4]
            i.grad = None; t.grad = None; log_s.grad = None
            fn(i, t, log_s.exp()).backward()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(k_steps):
            i, t = sets[k % 4]
            i.grad = None; t.grad = None; log_s.grad = None
            fn(i, t, log_s.exp()).backward()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / k_steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    ms_w = run(lb.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True, rank=rank,
                           world_size=world), steps)
    ms_1 = run(lb.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True), steps) \
        if world > 1 else ms_w
    tf_w = 6.0 * n * n_glob * DIM / (ms_w * 1e-3) / 1e12
    tf_1 = 6.0 * n * n * DIM / (ms_1 * 1e-3) / 1e12
    return {"rows_per_gpu": n, "global_batch": n_glob, "ms_per_step": ms_w,
            "samples_per_s": n_glob / (ms_w * 1e-3), "alg_tflops_per_gpu": tf_w,
            "single_gpu_same_box": {"global_batch": n, "ms_per_step": ms_1, "alg_tflops_per_gpu": tf_1},
            "efficiency_vs_single_gpu": tf_w / tf_1,
            "definition": "per-GPU credited FLOP/s (6 n N D / t) at n rows per GPU over W GPUs, divided by "
                          "the same GPU's at W = 1 (n = N), max over ranks"}


def run_ours(args):
    # NCCL prints its version banner on fd 1; keep stdout for the single JSON line
    sys.stdout.flush()
    saved_stdout_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        line, parity_ok = _run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout_fd, 1)
        os.close(saved_stdout_fd)
    if line is not None:
        print(json.dumps(line), flush=True)
    if not parity_ok:
        sys.stderr.write("PARITY FAILURE (see the \"parity\" key of the JSON line)\n")
        sys.exit(1)


def _run_ours(args):
    import torch.distributed as dist
    import latteclip_b200 as lb
    from latteclip_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    if args.gpus != world:
        if rank == 0:
            sys.stderr.write(f"note: --gpus {args.gpus} but WORLD_SIZE {world}; using {world}\n")
    n_loc = N_GLOBAL // world
    peaks = load_peaks()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        if os.environ.get("LATTE_BENCH_NO_SAMPLER") != "1":     # A/B knob: what the sampling itself costs
            sampler.start()                # nvidia-smi needs ~0.1 s before its first sample

    # rotating input sets: 4 x (I, T) x 32 MiB = 256 MiB of inputs at N=1, larger than the L2
    n_sets = 4
    host_sets = []
    dev_sets = []
    for sidx in range(n_sets):
        i, t = synth_shard(N_GLOBAL, DIM, rank, world, sidx)
        hi, ht = i.bfloat16().pin_memory(), t.bfloat16().pin_memory()
        host_sets.append((hi, ht))
        dev_sets.append((hi.to(dev).requires_grad_(True), ht.to(dev).requires_grad_(True)))
    log_s = torch.tensor(math.log(SCALE), device=dev, requires_grad=True)
    loss_fn = lb.ClipLoss(local_loss=True, gather_with_grad=True, cache_labels=True,
                          rank=rank, world_size=world)

    def step_resident(k):
        i, t = dev_sets[k % n_sets]
        i.grad = None; t.grad = None; log_s.grad = None
        loss = loss_fn(i, t, log_s.exp())
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # exactly W (>= 3) untimed warm-up steps, as the bench contract says (the first call also builds
    # the peer-memory state at N > 1)
    for k in range(max(args.warmup, 3)):
        step_resident(k)
        if k == 0:
            barrier()                      # workspaces / peer-memory state exist: the load is continuous from here
            sampler.load_start()
    barrier()

    # ---- timed region: exactly K steps, inputs resident in HBM ----------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    e0.record()
    for k in range(args.steps):
        loss = step_resident(k)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    # an identical second region, reported beside the first as information only (`repeat_ms_per_step`); the
    # clock sampler keeps running through it, so that a 60 ms region still gets two or three samples
    # of this very load at a sampling period that does not perturb it
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    r0.record()
    for k in range(args.steps):
        step_resident(k)
    r1.record()
    barrier()
    t_rep = torch.tensor([r0.elapsed_time(r1)], device=dev)
    if world > 1:
        dist.all_reduce(t_rep, op=dist.ReduceOp.MAX)
    repeat_ms_step = float(t_rep) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    t_ms = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_step = float(t_ms) / args.steps
    value = N_GLOBAL / (ms_step * 1e-3)
    last_loss = float(loss.detach())

    # ---- end to end: host (pinned) buffers -> public API -> loss back on the host ----------
    # Every step copies its own inputs from pinned host memory and reads its loss back; like a
    # pinned-memory data loader, the copy of step k+1 is issued on a copy stream while step k
    # computes (two staging slots, events both ways).  All K copies lie inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [(torch.empty(n_loc, DIM, dtype=torch.bfloat16, device=dev),
              torch.empty(n_loc, DIM, dtype=torch.bfloat16, device=dev)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    host_loss = torch.empty((), dtype=torch.float32).pin_memory()

    def prefetch(k):
        hi, ht = host_sets[k % n_sets]
        slot = k % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            slots[slot][0].copy_(hi, non_blocking=True)
            slots[slot][1].copy_(ht, non_blocking=True)
            ready[slot].record(copy_stream)

    def run_e2e(nsteps):
        cur = torch.cuda.current_stream(dev)
        for slot in range(2):
            consumed[slot].record(cur)
        prefetch(0)
        for k in range(nsteps):
            if k + 1 < nsteps:
                prefetch(k + 1)
            slot = k % 2
            cur.wait_event(ready[slot])
            i = slots[slot][0].detach().requires_grad_(True)
            t = slots[slot][1].detach().requires_grad_(True)
            log_s.grad = None
            loss = loss_fn(i, t, log_s.exp())
            loss.backward()
            host_loss.copy_(loss.detach(), non_blocking=True)
            consumed[slot].record(cur)

    run_e2e(3)
    barrier()
    e0.record()
    run_e2e(args.steps)
    e1.record()
    barrier()
    t_e = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = N_GLOBAL / (float(t_e) / args.steps * 1e-3)
    h2d = 2 * n_loc * DIM * 2
    d2h = 4

    # ---- parity of this very configuration (outside the timed regions) -----------------------
    pi, pt = dev_sets[1]
    parity = parity_check(loss_fn, pi, pt, log_s, rank, world, dev)

    # ---- roofline of the dominant kernel, timed live ------------------------------------------
    # The library records CUDA events on the launching stream around every kernel stage of
    # fwd + bwd (latte_clip_stage_times); the dominant stage is the stream-K gradient GEMM
    # (pair_gemm_kernel: dI = G.T_all and the text-side product G^T.I in ONE launch; at N > 1
    # the text side is an fp32 [N, D] partial that NCCL reduce-scatters).
    i, t = dev_sets[0]
    idet, tdet = i.detach(), t.detach()
    sc = torch.tensor(SCALE, device=dev)
    if world > 1:
        all_i = torch.empty(N_GLOBAL, DIM, dtype=torch.bfloat16, device=dev)
        all_t = torch.empty(N_GLOBAL, DIM, dtype=torch.bfloat16, device=dev)
        dist.all_gather_into_tensor(all_i, idet)
        dist.all_gather_into_tensor(all_t, tdet)
    else:
        all_i, all_t = idet, tdet
    off = rank * n_loc
    row, col, _ = _lib.clip_fwd(idet, tdet, all_i, all_t, off, sc)
    if world > 1:
        both = torch.empty(N_GLOBAL, 2, device=dev)
        dist.all_gather_into_tensor(both, torch.stack([row, col], dim=1))
        row_all, col_all = both[:, 0].contiguous(), both[:, 1].contiguous()
    else:
        row_all, col_all = row, col
    # N > 1 runs the one-sweep-per-rank flow (loss.py:_FusedClipLoss): time that flow.  The forward
    # stages come from the library's own event timing of step 1 of the rank flow; the backward
    # stages are timed on REAL backwards of the peer-memory flow (every rank calls
    # latte_clip_bwd_stage_times for the same generation), so the GEMM that is timed is the variant
    # that ran in the step: pair_gemm_kernel with the fused reduce-scatter into the peers'
    # accumulators.  Max over ranks.
    partial = world > 1 and _lib.rank_sweep_supported(torch.bfloat16, DIM)
    _lib.clip_stage_times(idet, tdet, all_i, all_t, off, sc, row_all, col_all, reps=2, partial=partial)
    stages = _lib.clip_stage_times(idet, tdet, all_i, all_t, off, sc, row_all, col_all, reps=8,
                                   partial=partial)
    gemm_variant = "pair_gemm_kernel (one launch: dI = G.T and dT = G^T.I)"
    if world > 1:
        from latteclip_b200 import loss as L
        op_i, _ = _lib.prep_features(idet, torch.bfloat16)
        op_t, _ = _lib.prep_features(tdet, torch.bfloat16)
        state = L._comm_state(n_loc, DIM, torch.float16, dev, None, world, rank)
        gemm_variant = "pair_gemm_kernel (dI = G.T local, dT = G^T.I as an fp32 partial for NCCL reduce_scatter)"
        if state is not None and L._bwd_sweeps(world) == 1:
            one = torch.ones(1, device=dev)
            acc = {k: 0.0 for k in _lib.STAGES}
            reps_real = 6
            for rep in range(2 + reps_real):
                slot = state.acquire()
                _lib.comm_push(slot.comm, op_t, None,
                               tensor_stride_bytes=slot.all_img.numel() * slot.all_img.element_size())
                r_all, rn_all, c_all, cn_all, _l, st_, _o = _lib.clip_fwd_rank(slot.comm, op_i, slot.all_txt,
                                                                               off, sc)
                sm = {}
                _lib.clip_bwd(op_i, op_t, None, slot.all_txt, off, sc, r_all, c_all, one, 1.0, True,
                              row_nll_all=rn_all, col_nll_all=cn_all, comm=slot.comm, lse_stats=st_,
                              stage_ms=sm)
                slot.release(signal=False)
                if rep >= 2:
                    for k in acc:
                        acc[k] += sm[k] / reps_real
            tb = torch.tensor([acc[k] for k in _lib.STAGES], device=dev)
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
            for k, v in zip(_lib.STAGES, tb.tolist()):
                if k.startswith("bwd"):
                    stages[k] = v
            gemm_variant = ("pair_gemm_kernel<peer TMA reduce> (dI = G.T local; dT = G^T.I added into the "
                            "owners' accumulators over NVLink from the epilogue: the fused reduce-scatter)")
    gemm_launches = 1
    gemm_ms = stages["bwd_gemm"]
    # algorithmic work of the stage: the two gradient GEMMs, 2 * n_loc * N * D FLOP each
    # (every FLOP of this kernel is credited work: it recomputes nothing)
    alg_flop_stage = 4.0 * n_loc * N_GLOBAL * DIM
    achieved_tf = alg_flop_stage / (gemm_ms * 1e-3) / 1e12
    step_tf = 6.0 * n_loc * N_GLOBAL * DIM / (ms_step * 1e-3) / 1e12
    roofline = {
        "bound": "tensor", "kernel": "pair_gemm_kernel", "variant": gemm_variant, "achieved": achieved_tf,
        "peak": peaks["burst"], "unit": "TFLOP/s", "frac": achieved_tf / peaks["burst"],
        "frac_of_sustained": achieved_tf / peaks["sustained"],
        "traffic": ncu_traffic_bytes() if world == 1 else None,
        "peak_source": peaks["source"] + ", burst bf16 figure (the stage is timed over 8 back-to-back "
                       "fwd+bwd repetitions, ~30 ms: too short for the power cap to pull clocks to the "
                       f"sustained level); sustained figure {peaks['sustained']:g} in frac_of_sustained",
        "launch_ms": gemm_ms / gemm_launches, "launches_per_step": gemm_launches,
        "alg_flop_per_launch": alg_flop_stage / gemm_launches,
        # G read once per product (fp16), the fp16 features, the fp32 accumulators
        "alg_bytes_per_launch": 2.0 * n_loc * N_GLOBAL * 2.0 + (N_GLOBAL + n_loc) * DIM * 2.0
                                + (N_GLOBAL + n_loc) * DIM * 4.0,
        "stage_ms": stages,
        "step_alg_tflops_per_gpu": step_tf,
        "step_frac_of_burst": step_tf / peaks["burst"],
        "step_frac_of_sustained": step_tf / peaks["sustained"],
        "executed_flop_per_step": 8.0 * n_loc * N_GLOBAL * DIM,
    }

    weak = weak_scaling_point(lb, rank, world, dev, args.steps)
    # the same at the headline per-GPU batch (32768 rows per GPU, global batch 32768 W): at W = 1 it
    # is the main measurement itself
    if world > 1:
        from latteclip_b200 import _lib as _l
        _l.clear_workspace_cache()
        torch.cuda.empty_cache()
        try:
            weak_big = weak_scaling_point(lb, rank, world, dev, min(args.steps, 10), rows=N_GLOBAL)
        except Exception as exc:           # an extra datum must not cost the line its headline numbers
            weak_big = {"rows_per_gpu": N_GLOBAL, "global_batch": N_GLOBAL * world,
                        "error": f"{type(exc).__name__}: {exc}"[:300]}
        _l.clear_workspace_cache()
        torch.cuda.empty_cache()
    else:
        weak_big = None

    # ---- prototype / pseudo-label kernels (HBM-bound rows of SURVEY 8a): achieved GB/s ------
    proto = None
    if rank == 0 and world == 1:
        proto = prototype_kernel_rates(dev, peaks)

    sig = None
    if rank == 0 and world == 1:
        _lib.clear_workspace_cache()
        sig = siglip_times(dev)
        _lib.clear_workspace_cache()

    head = head_big = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        head = latteclip_head_times(dev)
        head_big = latteclip_head_times(dev, batch=N_GLOBAL, reps=5, axis="row", cpu=False)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base, _ = cpu_reference_sample(4, 1)

    # kernels of ours per step (memset nodes and NCCL kernels not counted).  Forward: sweep,
    # finalize (rows + columns), gated fallback sweep, finish (gated merge + loss) (+ shard push
    # and column merge at N > 1); backward: LSE range + vectors, split-tile zeroing, sweep, GEMM,
    # split-tile cast, ds reduce.
    launches_per_step = 11 if world == 1 else 14
    bwd_mode = "one recompute sweep"
    if world > 1:
        from latteclip_b200.loss import _bwd_sweeps
        if _bwd_sweeps(world) == 2:
            launches_per_step += 4       # second sweep + its GEMM + the two fix-up kernels
            bwd_mode = "rows and columns recomputed per rank, no gradient exchange"
        else:
            bwd_mode = "one recompute sweep per rank, text gradient reduce-scattered inside the GEMM"
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": f"ClipLoss fwd+bwd (open_clip loss.py:120-130 + autograd), global batch "
                            f"{N_GLOBAL}, dim {DIM}, bf16 features, logit_scale {SCALE:g}, "
                            f"local_loss + gather_with_grad, {n_loc} rows per rank",
                "global_batch": N_GLOBAL, "dim": DIM, "rows_per_rank": n_loc,
                "l2": "rotating 4 input sets (256 MiB at N=1) larger than the 126 MiB L2",
                "loss": last_loss, "backward": bwd_mode,
                "repeat_ms_per_step": repeat_ms_step,
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks, "roofline": roofline, "parity": parity, "weak_scaling": weak,
        }
        if weak_big is not None:
            line["weak_scaling_32k_per_gpu"] = weak_big
        else:
            line["weak_scaling_32k_per_gpu"] = {
                "rows_per_gpu": N_GLOBAL, "global_batch": N_GLOBAL, "ms_per_step": ms_step,
                "efficiency_vs_single_gpu": 1.0,
                "definition": "the headline step itself at W = 1 (see the same key at N > 1)"}
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        if proto is not None:
            line["prototype_kernels"] = proto
        if sig is not None:
            line["siglip"] = sig
        if head is not None:
            line["latteclip_head"] = head
            line["latteclip_head_32k"] = head_big
    else:
        line = None
    if world > 1:
        dist.destroy_process_group()
    return line, parity["ok"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
