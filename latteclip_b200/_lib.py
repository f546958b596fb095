"""ctypes binding of liblatte_b200.so (C ABI in include/latte_b200.h).

PyTorch is used only for device memory and streams: every function here takes torch
tensors, checks them, and passes raw device pointers + sizes to the C ABI on the current
CUDA stream.  There is NO fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.
"""

from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from typing import Optional, Tuple

import torch

from ._build import SO_PATH, CSRC, SOURCES, build  # noqa: F401  (the nvcc recipe; setup.py uses it too)

F32, BF16, F16 = 0, 1, 2
LABEL_AXIS = {"row": 0, "quirk": 1}
_DTYPES = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}

_lib = None


def _declare(lib):
    c = ctypes
    vp, i64, i32, f32, sz = c.c_void_p, c.c_int64, c.c_int, c.c_float, c.c_size_t
    lib.latte_version.restype = c.c_char_p
    lib.latte_version.argtypes = []
    lib.latte_status_string.restype = c.c_char_p
    lib.latte_status_string.argtypes = [i32]
    lib.latte_device_info.argtypes = [c.POINTER(i32)] * 3
    lib.latte_clip_workspace_bytes.argtypes = [i64, i64, i64, i32, c.POINTER(sz)]
    lib.latte_clip_bwd_workspace_bytes.argtypes = [i64, i64, i64, i32, c.POINTER(sz)]
    lib.latte_clip_fwd.argtypes = [vp, i64, vp, i64, vp, i64, vp, i64, i32, i64, i64, i64, i64,
                                   vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.latte_clip_rank_sweep_supported.argtypes = [i32, i64]
    lib.latte_clip_fwd_rows.argtypes = [vp, i64, vp, i64, i32, i64, i64, i64, i64, vp,
                                        vp, vp, vp, vp, vp, sz, vp]
    lib.latte_clip_fwd_cols_workspace_bytes.argtypes = [i64, i64, i32, c.POINTER(sz)]
    lib.latte_clip_fwd_cols.argtypes = [vp, i64, i32, vp, i64, vp, i64, i32, i64, i64, i64, i64,
                                        vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.latte_clip_bwd.argtypes = [vp, i64, vp, i64, vp, i64, vp, i64, i32, i64, i64, i64, i64,
                                   vp, vp, vp, vp, vp, vp, vp, f32, i32, vp, vp, i32, i64, vp, vp, i32,
                                   vp, vp, sz, vp]    # ..., d_txt_partial, comm, phases, d_scale, ws, bytes, stream
    lib.latte_clip_bwd_stage_times.argtypes = lib.latte_clip_bwd.argtypes + [c.POINTER(f32)]
    lib.latte_clip_stage_times.argtypes = [vp, i64, vp, i64, vp, i64, vp, i64, i32, i64, i64, i64, i64,
                                           vp, vp, vp, vp, vp, vp, vp, f32, i32, vp, vp, i32, i64,
                                           vp, vp, vp, sz, vp, sz, vp, i32, c.POINTER(f32)]
    lib.latte_prep_features.argtypes = [vp, i64, i32, i64, i64, i32, i32, vp, i64, vp, vp]
    lib.latte_normalize_bwd.argtypes = [vp, i64, i32, vp, i64, i32, vp, i64, i64, vp, i64, vp]
    lib.latte_comm_push.argtypes = [vp, vp, vp, i64, i64, vp]
    lib.latte_comm_release.argtypes = [vp, vp]
    lib.latte_clip_fwd_rank_workspace_bytes.argtypes = [i64, i64, i64, i32, c.POINTER(sz)]
    lib.latte_clip_fwd_rank.argtypes = [vp, vp, i64, vp, i64, i32, i64, i64, i64, i64, vp,
                                        vp, vp, vp, vp, vp, vp, i32, vp, sz, vp]
    lib.latte_siglip_supported.argtypes = [i32, i64]
    lib.latte_siglip_workspace_bytes.argtypes = [i64, i64, i64, i32, i32, i32, c.POINTER(sz)]
    lib.latte_siglip_fwd.argtypes = [vp, i64, vp, i64, i32, i64, i64, i64, i64, vp, vp, vp, vp, sz, vp]
    lib.latte_siglip_bwd.argtypes = [vp, i64, vp, i64, i32, i64, i64, i64, i64, vp, vp, vp, vp, vp,
                                     i32, i64, vp, vp, i32, vp, vp, vp, sz, vp]
    lib.latte_normalize_rows.argtypes = [vp, i64, vp, i64, i64, i64, vp]
    lib.latte_nxc_workspace_bytes.argtypes = [i32, i32, i64, i64, i64, c.POINTER(sz)]
    lib.latte_nxc_argmax_margin.argtypes = [vp, i64, i32, vp, i64, i64, vp, i64, i64, f32,
                                            vp, vp, vp, vp, sz, vp]
    lib.latte_nxc_topk.argtypes = [vp, i64, i32, i64, i64, vp, i64, i64, f32, i32, vp, vp, vp, sz, vp]
    lib.latte_seg_workspace_bytes.argtypes = [i64, i64, i64, c.POINTER(sz)]
    lib.latte_nxc_planes_bytes.argtypes = [i64, i64, c.POINTER(sz)]
    lib.latte_nxc_split_prototypes.argtypes = [vp, i64, i64, i64, i32, vp, vp, i64, vp]
    lib.latte_nxc_multi.argtypes = [vp, i32, vp]
    lib.latte_mix_ema_fwd.argtypes = [vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, vp, vp, vp, vp,
                                      f32, i32, i32, i64, i64, i64, vp, vp, i64, vp]
    lib.latte_mix_ema_bwd.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, vp, f32, i32, i32, i64,
                                      i64, i64, vp, i64, vp, vp, i64, vp, i64, vp, sz, vp]
    lib.latte_bank_accumulate.argtypes = [vp, vp, i64, i32, vp, vp, i64, i64, i64, vp, i64, vp,
                                          vp, sz, vp]
    lib.latte_bank_finalize.argtypes = [vp, i64, vp, vp, i64, i64, i64, vp]
    lib.latte_distill_aux_bytes.argtypes = [c.POINTER(sz)]
    lib.latte_distill_products.argtypes = [vp, i64, vp, i64, vp, i64, vp, i64, i64, i64, vp, vp, vp,
                                           vp, vp, i64, vp, sz, vp]
    lib.latte_distill_loss.argtypes = [vp, vp, i64, vp, i64, vp, i64, i64, vp, vp, vp, vp, sz, vp]
    lib.latte_distill_bwd_combine.argtypes = [vp, vp, vp, vp, i64, vp, i64, i64, i64, vp, vp, vp, vp, vp,
                                              i32, i64, vp, vp, sz, vp]
    for name in EXPORTS:
        if name not in ("latte_version", "latte_status_string"):
            getattr(lib, name).restype = i32


EXPORTS = [
    "latte_version", "latte_status_string", "latte_device_info", "latte_clip_workspace_bytes",
    "latte_clip_bwd_workspace_bytes", "latte_clip_stage_times", "latte_clip_bwd_stage_times", "latte_clip_rank_sweep_supported",
    "latte_clip_fwd_rows", "latte_clip_fwd_cols_workspace_bytes", "latte_clip_fwd_cols",
    "latte_prep_features", "latte_normalize_bwd", "latte_comm_push", "latte_comm_release", "latte_clip_fwd_rank_workspace_bytes",
    "latte_clip_fwd_rank", "latte_siglip_supported", "latte_siglip_workspace_bytes",
    "latte_siglip_fwd", "latte_siglip_bwd",
    "latte_clip_fwd", "latte_clip_bwd", "latte_normalize_rows", "latte_nxc_workspace_bytes",
    "latte_nxc_argmax_margin", "latte_nxc_topk", "latte_seg_workspace_bytes", "latte_nxc_planes_bytes",
    "latte_nxc_split_prototypes", "latte_nxc_multi", "latte_mix_ema_fwd", "latte_mix_ema_bwd", "latte_bank_accumulate",
    "latte_bank_finalize", "latte_distill_aux_bytes", "latte_distill_products", "latte_distill_loss",
    "latte_distill_bwd_combine",
]


def load():
    """Load the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} is missing: build the CUDA extension first "
                "(python -c 'import __graft_entry__ as g; g.build()' or python -m latteclip_b200.build). "
                "latteclip_b200 has no CPU / PyTorch fallback.")
        lib = ctypes.CDLL(SO_PATH)
        _declare(lib)
        _lib = lib
    return _lib


def version() -> str:
    return load().latte_version().decode()


def _check(status: int, what: str):
    if status != 0:
        msg = load().latte_status_string(status).decode()
        raise RuntimeError(f"{what} failed: {msg} (status {status})")


def _stream(t: torch.Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _dt(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise RuntimeError(f"unsupported dtype {t.dtype}") from None


def _rows(t: torch.Tensor, name: str) -> torch.Tensor:
    """2-D, unit stride along the feature axis (row stride is passed to the kernels)."""
    if t.dim() != 2:
        raise RuntimeError(f"{name} must be 2-D, got shape {tuple(t.shape)}")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: latteclip_b200 has no CPU path")
    if t.stride(1) != 1 or t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t


def _vec(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    return t.detach().to(dtype).contiguous()


def _scalar_f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).reshape(1).contiguous()


def _clip_ws_bytes(n_loc: int, n_all: int, dim: int, dtype: int, bwd: bool = False) -> int:
    nbytes = ctypes.c_size_t(0)
    fn = load().latte_clip_bwd_workspace_bytes if bwd else load().latte_clip_workspace_bytes
    _check(fn(n_loc, n_all, dim, dtype, ctypes.byref(nbytes)), "latte_clip_workspace_bytes")
    return nbytes.value


def _aligned_ptr(ws: torch.Tensor) -> Tuple[ctypes.c_void_p, int]:
    p = ws.data_ptr()
    off = (-p) % 256
    return ctypes.c_void_p(p + off), ws.numel() - off


# Scratch for the C ABI's caller-provided workspaces: ONE growable buffer per (kind, device,
# stream).  The kernels of one call are stream-ordered before the next call's on the same
# stream, so a kind's buffer can be reused by every call of that kind; a larger request
# replaces the buffer (the old one goes back to the caching allocator, which keeps it alive for
# the work already enqueued on that stream).  Sizes come from the *_workspace_bytes functions
# -- nothing is allocated to measure.  The backward buffer holds the fp16 gradient weights
# G [n_loc, N] (2 GiB at N = 32768 on one GPU); clear_workspace_cache() drops everything.
_WS_CACHE = {}


def _scratch(kind: str, nbytes: int, device) -> torch.Tensor:
    k = (kind, device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _WS_CACHE.get(k)
    need = int(nbytes) + 256
    if ws is None or ws.numel() < need:
        _WS_CACHE.pop(k, None)
        ws = None                      # release the old buffer before allocating the new one
        ws = torch.empty(need, dtype=torch.uint8, device=device)
        _WS_CACHE[k] = ws
    return ws


def clear_workspace_cache():
    """Drop the cached scratch buffers (e.g. 2 GiB of gradient weights at batch 32768)."""
    _WS_CACHE.clear()


def workspace_cache_bytes() -> int:
    return sum(int(t.numel()) for t in _WS_CACHE.values())


# ------------------------------------------------------------------------------ operands
def prep_features(x: torch.Tensor, compute_dtype: torch.dtype, normalize: bool = False):
    """Fused (normalise +) cast of a feature matrix into the fp16 tensor-core operand:
    ``fp16(round_to(compute_dtype, normalize(x)))`` -> (operand fp16 [n, dim], inv_norm fp32 [n] or
    None).  fp16 inputs with fp16 compute and no normalisation are returned as they are (no kernel)."""
    x = _rows(x.detach(), "features")
    if x.dtype == torch.float16 and compute_dtype == torch.float16 and not normalize and \
            x.stride(0) % 8 == 0 and x.data_ptr() % 16 == 0:
        return x, None
    if x.stride(0) % 8 != 0 or x.data_ptr() % 16 != 0:
        x = x.contiguous()
    n, dim = x.shape
    out = torch.empty(n, dim, dtype=torch.float16, device=x.device)
    inv = torch.empty(n, dtype=torch.float32, device=x.device) if normalize else None
    with torch.cuda.device(x.device):
        _check(load().latte_prep_features(_ptr(x), x.stride(0), _dt(x), n, dim, int(bool(normalize)),
                                          _DTYPES[compute_dtype], _ptr(out), dim, _ptr(inv), _stream(x)),
               "latte_prep_features")
    return out, inv


def normalize_bwd(g: torch.Tensor, x: torch.Tensor, inv_norm: torch.Tensor) -> torch.Tensor:
    """d_x of ``F.normalize(x, dim=-1)`` given the gradient g of the normalised rows."""
    g, x = _rows(g.detach(), "grad"), _rows(x.detach(), "features")
    out = torch.empty(g.shape, dtype=x.dtype, device=g.device)
    with torch.cuda.device(g.device):
        _check(load().latte_normalize_bwd(_ptr(g), g.stride(0), _dt(g), _ptr(x), x.stride(0), _dt(x),
                                          _ptr(inv_norm), g.shape[0], g.shape[1], _ptr(out), out.stride(0),
                                          _stream(g)),
               "latte_normalize_bwd")
    return out


# ------------------------------------------------------------------------------ ClipLoss
def clip_fwd(img_loc, txt_loc, img_all, txt_all, label_offset: int, logit_scale, with_nll: bool = False,
             with_stats: bool = False):
    """-> (row_lse[n_loc], col_lse[n_loc], loss[1]) fp32 device tensors; with_nll appends
    (row_nll[n_loc], col_nll[n_loc]), the per-sample loss terms lse - label logit; with_stats
    appends stats[4] (LSE min / max, largest nll: the backward's scaling inputs)."""
    lib = load()
    img_loc, txt_loc = _rows(img_loc, "image_features"), _rows(txt_loc, "text_features")
    img_all, txt_all = _rows(img_all, "all_image_features"), _rows(txt_all, "all_text_features")
    dt = _dt(img_loc)
    if not (_dt(txt_loc) == _dt(img_all) == _dt(txt_all) == dt):
        raise RuntimeError("clip_fwd: all feature tensors must share one dtype")
    n_loc, dim = img_loc.shape
    n_all = img_all.shape[0]
    dev = img_loc.device
    s = _scalar_f32(logit_scale)
    row_lse = torch.empty(n_loc, dtype=torch.float32, device=dev)
    col_lse = torch.empty(n_loc, dtype=torch.float32, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    row_nll = torch.empty(n_loc, dtype=torch.float32, device=dev) if with_nll else None
    col_nll = torch.empty(n_loc, dtype=torch.float32, device=dev) if with_nll else None
    stats = torch.empty(4, dtype=torch.float32, device=dev) if with_stats else None
    ws = _scratch("fwd", _clip_ws_bytes(n_loc, n_all, dim, dt), dev)
    wp, wn = _aligned_ptr(ws)
    with torch.cuda.device(dev):
        _check(lib.latte_clip_fwd(_ptr(img_loc), img_loc.stride(0), _ptr(txt_loc), txt_loc.stride(0),
                                  _ptr(img_all), img_all.stride(0), _ptr(txt_all), txt_all.stride(0),
                                  dt, n_loc, n_all, dim, label_offset, _ptr(s), _ptr(row_lse),
                                  _ptr(col_lse), _ptr(row_nll), _ptr(col_nll), _ptr(loss), _ptr(stats),
                                  wp, wn, _stream(img_loc)),
               "latte_clip_fwd")
    out = (row_lse, col_lse, loss)
    if with_nll:
        out = out + (row_nll, col_nll)
    if with_stats:
        out = out + (stats,)
    return out


COMM_MAX_RANKS = 8
COMM_FLAG_INTS = 64


class LatteComm(ctypes.Structure):
    """latte_comm_t of include/latte_b200.h: one slot of the peer-memory exchange."""
    _fields_ = [("rank", ctypes.c_int32), ("world", ctypes.c_int32), ("gen", ctypes.c_int32),
                ("reserved", ctypes.c_int32),
                ("gather", ctypes.c_void_p * COMM_MAX_RANKS),
                ("payload", ctypes.c_void_p * COMM_MAX_RANKS),
                ("acc", ctypes.c_void_p * COMM_MAX_RANKS),
                ("flags", ctypes.c_void_p * COMM_MAX_RANKS),
                ("payload_stride", ctypes.c_int64)]


def make_comm(rank: int, world: int, gen: int, gather_ptrs, payload_ptrs, acc_ptrs, flag_ptrs,
              payload_stride: int) -> LatteComm:
    c = LatteComm()
    c.rank, c.world, c.gen, c.reserved = int(rank), int(world), int(gen), 0
    for w in range(world):
        c.gather[w] = int(gather_ptrs[w])
        c.payload[w] = int(payload_ptrs[w])
        c.acc[w] = int(acc_ptrs[w])
        c.flags[w] = int(flag_ptrs[w]) if flag_ptrs is not None else None
    c.payload_stride = int(payload_stride)
    return c


def comm_push(comm: LatteComm, txt_shard, img_shard=None, tensor_stride_bytes: int = 0):
    """The feature all-gather as a store kernel into every rank's gather buffer (text matrix at
    ``tensor_stride_bytes``), publishing landed flags; waits for the slot's credits first."""
    txt_shard = _rows(txt_shard, "text_features").contiguous()
    if img_shard is not None:
        img_shard = _rows(img_shard, "image_features").contiguous()
    nbytes = txt_shard.numel() * txt_shard.element_size()
    with torch.cuda.device(txt_shard.device):
        _check(load().latte_comm_push(ctypes.byref(comm), _ptr(txt_shard), _ptr(img_shard), nbytes,
                                      int(tensor_stride_bytes), _stream(txt_shard)),
               "latte_comm_push")


def comm_release(comm: LatteComm, device):
    with torch.cuda.device(device):
        _check(load().latte_comm_release(ctypes.byref(comm),
                                         ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)),
               "latte_comm_release")


def clip_fwd_rank(comm: LatteComm, img_loc, txt_all, label_offset: int, logit_scale, phases: int = 7,
                  out=None):
    """Multi-rank forward over peer memory (after comm_push of this generation) ->
    (row_lse_all, row_nll_all, col_lse_all, col_nll_all [n_all each], loss[1], stats[4])."""
    lib = load()
    img_loc, txt_all = _rows(img_loc, "image_features"), _rows(txt_all, "all_text_features")
    dt = _dt(img_loc)
    n_loc, dim = img_loc.shape
    n_all = txt_all.shape[0]
    dev = img_loc.device
    s = _scalar_f32(logit_scale)
    if out is None:
        vec = torch.empty(4, n_all, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        stats = torch.empty(4, dtype=torch.float32, device=dev)
    else:
        vec, loss, stats = out
    optr = vec.data_ptr()
    vecs = [ctypes.c_void_p(optr + 4 * n_all * k) for k in range(4)]
    nbytes = ctypes.c_size_t(0)
    _check(lib.latte_clip_fwd_rank_workspace_bytes(n_loc, n_all, dim, dt, ctypes.byref(nbytes)),
           "latte_clip_fwd_rank_workspace_bytes")
    ws = _scratch("fwd_rank" if phases == 7 else f"fwd_rank{comm.rank}", nbytes.value, dev)
    wp, wn = _aligned_ptr(ws)
    with torch.cuda.device(dev):
        _check(lib.latte_clip_fwd_rank(ctypes.byref(comm), _ptr(img_loc), img_loc.stride(0),
                                       _ptr(txt_all), txt_all.stride(0), dt, n_loc, n_all, dim,
                                       int(label_offset), _ptr(s), vecs[0], vecs[1], vecs[2], vecs[3],
                                       _ptr(loss), _ptr(stats), int(phases), wp, wn, _stream(img_loc)),
               "latte_clip_fwd_rank")
    return vec[0], vec[1], vec[2], vec[3], loss, stats, (vec, loss, stats)


def rank_sweep_supported(dtype: torch.dtype, dim: int) -> bool:
    """True when the one-sweep-per-rank multi-rank path (clip_fwd_rows / clip_fwd_cols /
    clip_bwd(partial=True)) can run these features."""
    if dtype not in _DTYPES:
        return False
    return bool(load().latte_clip_rank_sweep_supported(_DTYPES[dtype], dim))


def clip_fwd_rows(img_loc, txt_all, label_offset: int, logit_scale):
    """Step 1 of the multi-rank forward -> packed fp32 payload [2 n_all + 3 n_loc]:
    col_ml [n_all, 2] | row_lse | row_nll | label_logit  (what the ranks all-gather)."""
    lib = load()
    img_loc, txt_all = _rows(img_loc, "image_features"), _rows(txt_all, "all_text_features")
    dt = _dt(img_loc)
    n_loc, dim = img_loc.shape
    n_all = txt_all.shape[0]
    dev = img_loc.device
    s = _scalar_f32(logit_scale)
    payload = torch.empty(2 * n_all + 3 * n_loc, dtype=torch.float32, device=dev)
    base = payload.data_ptr()
    col_ml = ctypes.c_void_p(base)
    row_lse = ctypes.c_void_p(base + 8 * n_all)
    row_nll = ctypes.c_void_p(base + 8 * n_all + 4 * n_loc)
    label_logit = ctypes.c_void_p(base + 8 * n_all + 8 * n_loc)

    ws = _scratch("fwd", _clip_ws_bytes(n_loc, n_all, dim, dt), dev)
    wp, wn = _aligned_ptr(ws)
    with torch.cuda.device(dev):
        _check(lib.latte_clip_fwd_rows(_ptr(img_loc), img_loc.stride(0), _ptr(txt_all), txt_all.stride(0),
                                       dt, n_loc, n_all, dim, label_offset, _ptr(s), row_lse,
                                       row_nll, label_logit, col_ml, wp, wn, _stream(img_loc)),
               "latte_clip_fwd_rows")
    return payload


def clip_fwd_cols(gathered, img_all, txt_all, n_loc: int, label_offset: int, logit_scale):
    """Step 2 of the multi-rank forward.  ``gathered``: the all-gathered payloads
    [world, 2 n_all + 3 n_loc] -> (row_lse_all, row_nll_all, col_lse_all, col_nll_all) [n_all]
    each and loss[1]."""
    lib = load()
    img_all, txt_all = _rows(img_all, "all_image_features"), _rows(txt_all, "all_text_features")
    dt = _dt(img_all)
    n_all, dim = img_all.shape
    dev = img_all.device
    if gathered.dim() != 2 or gathered.dtype != torch.float32 or gathered.stride(1) != 1:
        raise RuntimeError("clip_fwd_cols: gathered must be a 2-D fp32 tensor")
    world = gathered.shape[0]
    if gathered.shape[1] != 2 * n_all + 3 * n_loc or n_all != n_loc * world:
        raise RuntimeError("clip_fwd_cols: gathered payload has the wrong size")
    s = _scalar_f32(logit_scale)
    out = torch.empty(4, n_all, dtype=torch.float32, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    stats = torch.empty(4, dtype=torch.float32, device=dev)
    optr = out.data_ptr()
    vecs = [ctypes.c_void_p(optr + 4 * n_all * k) for k in range(4)]

    nbytes = ctypes.c_size_t(0)
    _check(lib.latte_clip_fwd_cols_workspace_bytes(n_all, dim, dt, ctypes.byref(nbytes)),
           "latte_clip_fwd_cols_workspace_bytes")
    ws = _scratch("cols", nbytes.value, dev)
    wp, wn = _aligned_ptr(ws)
    with torch.cuda.device(dev):
        _check(lib.latte_clip_fwd_cols(_ptr(gathered), gathered.stride(0), world,
                                       _ptr(img_all), img_all.stride(0), _ptr(txt_all), txt_all.stride(0),
                                       dt, n_loc, n_all, dim, label_offset, _ptr(s), vecs[0], vecs[1],
                                       vecs[2], vecs[3], _ptr(loss), _ptr(stats), wp, wn,
                                       _stream(img_all)),
               "latte_clip_fwd_cols")
    return out[0], out[1], out[2], out[3], loss, stats


def clip_bwd(img_loc, txt_loc, img_all, txt_all, label_offset: int, logit_scale,
             row_lse_all, col_lse_all, grad_loss, grad_mult: float, cross_terms: bool,
             grad_dtype=None, row_nll_all=None, col_nll_all=None, partial: bool = False,
             comm: Optional[LatteComm] = None, phases: int = 3, lse_stats=None, out=None,
             stage_ms: Optional[dict] = None):
    """-> (d_img[n_loc, dim], d_txt[n_loc, dim], d_scale[1] fp32).  The feature gradients
    come back in ``grad_dtype`` (default: the feature dtype, what autograd needs).
    ``partial=True`` (one-sweep multi-rank mode): the second result is instead the fp32
    partial [n_all, dim] of the text gradient over ALL columns, to be reduce-scattered.
    ``comm`` (one-sweep mode, fused reduce-scatter over peer memory): the text-side product is added
    straight into the owners' accumulators from the GEMM epilogue and the call's last kernel turns
    this rank's accumulator into d_txt once every rank's adds have landed.  In the one-sweep modes
    ``img_all`` may be None (only this rank's images are read).  ``phases`` / ``out`` let a
    single-process test drive the ranks phase by phase (out = the result tensors of phase 1)."""
    lib = load()
    img_loc, txt_loc = _rows(img_loc, "image_features"), _rows(txt_loc, "text_features")
    txt_all = _rows(txt_all, "all_text_features")
    if img_all is not None:
        img_all = _rows(img_all, "all_image_features")
    elif not (partial or comm is not None):
        raise RuntimeError("clip_bwd: img_all is required outside the one-sweep modes")
    dt = _dt(img_loc)
    n_loc, dim = img_loc.shape
    n_all = txt_all.shape[0]
    dev = img_loc.device
    s = _scalar_f32(logit_scale)
    g = _scalar_f32(grad_loss)
    row_lse_all = _vec(row_lse_all, torch.float32, "row_lse")
    col_lse_all = _vec(col_lse_all, torch.float32, "col_lse")
    if row_lse_all.numel() != n_all or col_lse_all.numel() != n_all:
        raise RuntimeError("clip_bwd: LSE vectors must have n_all entries")
    gdt = img_loc.dtype if grad_dtype is None else grad_dtype
    if (row_nll_all is None) != (col_nll_all is None):
        raise RuntimeError("clip_bwd: pass both nll vectors or neither")
    if row_nll_all is not None:
        row_nll_all = _vec(row_nll_all, torch.float32, "row_nll")
        col_nll_all = _vec(col_nll_all, torch.float32, "col_nll")
        if row_nll_all.numel() != n_all or col_nll_all.numel() != n_all:
            raise RuntimeError("clip_bwd: nll vectors must have n_all entries")
    if out is None:
        d_img = torch.empty(n_loc, dim, dtype=gdt, device=dev)
        d_txt = None if partial else torch.empty(n_loc, dim, dtype=gdt, device=dev)
        d_part = torch.empty(n_all, dim, dtype=torch.float32, device=dev) if partial else None
        d_scale = torch.empty(1, dtype=torch.float32, device=dev)
    else:
        d_img, d_txt, d_part, d_scale = out
    ws = _scratch("bwd" if phases == 3 else f"bwd{0 if comm is None else comm.rank}",
                  _clip_ws_bytes(n_loc, n_all, dim, dt, bwd=True), dev)
    wp, wn = _aligned_ptr(ws)
    args = (_ptr(img_loc), img_loc.stride(0), _ptr(txt_loc), txt_loc.stride(0),
            _ptr(img_all), 0 if img_all is None else img_all.stride(0),
            _ptr(txt_all), txt_all.stride(0),
            dt, n_loc, n_all, dim, label_offset, _ptr(s), _ptr(row_lse_all),
            _ptr(col_lse_all), _ptr(row_nll_all), _ptr(col_nll_all),
            _ptr(lse_stats), _ptr(g),
            float(grad_mult), int(bool(cross_terms)),
            _ptr(d_img), _ptr(d_txt), _DTYPES[gdt], dim, _ptr(d_part),
            ctypes.byref(comm) if comm is not None else None, int(phases),
            _ptr(d_scale), wp, wn, _stream(img_loc))
    with torch.cuda.device(dev):
        if stage_ms is not None:
            # the same call with CUDA events around its stages (synchronises the stream)
            buf = (ctypes.c_float * len(STAGES))()
            _check(lib.latte_clip_bwd_stage_times(*args, buf), "latte_clip_bwd_stage_times")
            stage_ms.update({name: float(buf[k]) for k, name in enumerate(STAGES)})
        else:
            _check(lib.latte_clip_bwd(*args), "latte_clip_bwd")
    if out is not None or phases != 3:
        return d_img, (d_part if partial else d_txt), d_scale, (d_img, d_txt, d_part, d_scale)
    return d_img, (d_part if partial else d_txt), d_scale


# ------------------------------------------------------------------------------ DistillClipLoss
def _distill_aux(dev):
    need = ctypes.c_size_t()
    _check(load().latte_distill_aux_bytes(ctypes.byref(need)), "latte_distill_aux_bytes")
    ws = _scratch("distill_aux", need.value, dev)
    return _aligned_ptr(ws)


def distill_products(sweep_img, sweep_txt, gemm_img, gemm_txt, logit_scale, row_lse, col_lse):
    """W = softmax_rows(S) + softmax_cols(S) of S = s * sweep_img @ sweep_txt.T (never stored), multiplied
    into the other operand pair: -> (W @ gemm_txt, W.T @ gemm_img), fp32 [n, dim].  All four matrices
    fp16 [n, dim]; row_lse / col_lse are the forward's LSE vectors of S (latte_clip_fwd)."""
    mats = [_rows(x.detach(), "features") for x in (sweep_img, sweep_txt, gemm_img, gemm_txt)]
    if any(m.dtype != torch.float16 for m in mats):
        raise RuntimeError("distill_products: operands must be fp16 (latte_prep_features)")
    n, dim = mats[0].shape
    dev = mats[0].device
    out_i = torch.empty(n, dim, dtype=torch.float32, device=dev)
    out_t = torch.empty(n, dim, dtype=torch.float32, device=dev)
    s = _scalar_f32(logit_scale)
    ws = _scratch("bwd", _clip_ws_bytes(n, n, dim, _DTYPES[torch.float16], bwd=True), dev)
    wp, wn = _aligned_ptr(ws)
    with torch.cuda.device(dev):
        _check(load().latte_distill_products(
            _ptr(mats[0]), mats[0].stride(0), _ptr(mats[1]), mats[1].stride(0), _ptr(mats[2]),
            mats[2].stride(0), _ptr(mats[3]), mats[3].stride(0), n, dim, _ptr(s),
            _ptr(_vec(row_lse, torch.float32, "row_lse")), _ptr(_vec(col_lse, torch.float32, "col_lse")),
            _ptr(out_i), _ptr(out_t), dim, wp, wn, _stream(mats[0])), "latte_distill_products")
    return out_i, out_t


def distill_loss(row_lse, col_lse, img, teacher_prod, logit_scale):
    """-> (loss[1], dot[1]): loss = (sum row_lse + sum col_lse - s * <img, teacher_prod>) / (2 n)."""
    img = _rows(img.detach(), "image_features")
    n, dim = img.shape
    dev = img.device
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    dot = torch.empty(1, dtype=torch.float32, device=dev)
    s = _scalar_f32(logit_scale)
    ap, an = _distill_aux(dev)
    with torch.cuda.device(dev):
        _check(load().latte_distill_loss(_ptr(_vec(row_lse, torch.float32, "row_lse")),
                                         _ptr(_vec(col_lse, torch.float32, "col_lse")), n, _ptr(img),
                                         img.stride(0), _ptr(teacher_prod), teacher_prod.stride(0), dim,
                                         _ptr(s), _ptr(loss), _ptr(dot), ap, an, _stream(img)),
               "latte_distill_loss")
    return loss, dot


def distill_bwd_combine(a_s, a_t, b_s, b_t, img, logit_scale, grad_loss, dot_t, grad_dtype):
    """-> (d_img, d_txt [n, dim] in grad_dtype, d_scale[1]) from the student / teacher products."""
    img = _rows(img.detach(), "image_features")
    n, dim = img.shape
    dev = img.device
    d_img = torch.empty(n, dim, dtype=grad_dtype, device=dev)
    d_txt = torch.empty(n, dim, dtype=grad_dtype, device=dev)
    d_scale = torch.empty(1, dtype=torch.float32, device=dev)
    s, g = _scalar_f32(logit_scale), _scalar_f32(grad_loss)
    ap, an = _distill_aux(dev)
    with torch.cuda.device(dev):
        _check(load().latte_distill_bwd_combine(_ptr(a_s), _ptr(a_t), _ptr(b_s), _ptr(b_t), a_s.stride(0),
                                                _ptr(img), img.stride(0), n, dim, _ptr(s), _ptr(g),
                                                _ptr(dot_t), _ptr(d_img), _ptr(d_txt), _DTYPES[grad_dtype],
                                                dim, _ptr(d_scale), ap, an, _stream(img)),
               "latte_distill_bwd_combine")
    return d_img, d_txt, d_scale


def siglip_supported(dtype: torch.dtype, dim: int) -> bool:
    """True when the SigLIP kernels take features of this dtype / width (16-bit, dim <= 768, % 8)."""
    if dtype not in _DTYPES:
        return False
    return bool(load().latte_siglip_supported(_DTYPES[dtype], int(dim)))


def _siglip_workspace(n_loc: int, n_all: int, dim: int, dt: int, dev, bwd: bool, own: bool):
    need = ctypes.c_size_t()
    _check(load().latte_siglip_workspace_bytes(n_loc, n_all, dim, dt, int(bwd), int(own),
                                               ctypes.byref(need)), "latte_siglip_workspace_bytes")
    return _scratch("siglip_bwd" if bwd else "siglip_fwd", need.value + 256, dev)


def siglip_fwd(img_loc, txt_all, label_offset: int, logit_scale, logit_bias=None):
    """-> loss[1] fp32 of this rank: sum over own image rows x ALL text columns of
    -logsigmoid(label * (s <i, t> + b)) / n_loc  (open_clip loss.py:509-519)."""
    lib = load()
    img_loc, txt_all = _rows(img_loc, "image_features"), _rows(txt_all, "all_text_features")
    if img_loc.dtype != txt_all.dtype or img_loc.shape[1] != txt_all.shape[1]:
        raise RuntimeError("siglip_fwd: image and text features must share dtype and width")
    dt = _dt(img_loc)
    n_loc, dim = img_loc.shape
    n_all = txt_all.shape[0]
    dev = img_loc.device
    s = _scalar_f32(logit_scale)
    b = _scalar_f32(logit_bias) if logit_bias is not None else None
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    ws = _siglip_workspace(n_loc, n_all, dim, dt, dev, False, False)
    wp, wn = _aligned_ptr(ws)
    with torch.cuda.device(dev):
        _check(lib.latte_siglip_fwd(_ptr(img_loc), img_loc.stride(0), _ptr(txt_all), txt_all.stride(0),
                                    dt, n_loc, n_all, dim, int(label_offset), _ptr(s), _ptr(b),
                                    _ptr(loss), wp, wn, _stream(img_loc)), "latte_siglip_fwd")
    return loss


def siglip_bwd(img_loc, txt_all, label_offset: int, logit_scale, logit_bias, grad_loss,
               grad_dtype=None, partial: bool = False, peer_ptrs=None):
    """-> (d_img [n_loc, dim], d_txt, d_scale[1], d_bias[1]).  d_txt is [n_all, dim] in
    ``grad_dtype`` when this rank holds every text row (n_loc == n_all); with ``partial=True`` the
    fp32 partial [n_all, dim] for the caller's reduce-scatter; with ``peer_ptrs`` None (the product
    was added into the peers' accumulators)."""
    lib = load()
    img_loc, txt_all = _rows(img_loc, "image_features"), _rows(txt_all, "all_text_features")
    dt = _dt(img_loc)
    n_loc, dim = img_loc.shape
    n_all = txt_all.shape[0]
    dev = img_loc.device
    s = _scalar_f32(logit_scale)
    b = _scalar_f32(logit_bias) if logit_bias is not None else None
    g = _scalar_f32(grad_loss)
    gdt = img_loc.dtype if grad_dtype is None else grad_dtype
    fused = peer_ptrs is not None
    own = not (partial or fused)
    if own and n_loc != n_all:
        raise RuntimeError("siglip_bwd: d_txt needs partial=True or peer_ptrs when n_loc < n_all")
    d_img = torch.empty(n_loc, dim, dtype=gdt, device=dev)
    d_txt = torch.empty(n_all, dim, dtype=gdt, device=dev) if own else None
    d_part = torch.empty(n_all, dim, dtype=torch.float32, device=dev) if partial else None
    peers = (ctypes.c_void_p * len(peer_ptrs))(*peer_ptrs) if fused else None
    d_scale = torch.empty(1, dtype=torch.float32, device=dev)
    d_bias = torch.empty(1, dtype=torch.float32, device=dev)
    ws = _siglip_workspace(n_loc, n_all, dim, dt, dev, True, own)
    wp, wn = _aligned_ptr(ws)
    with torch.cuda.device(dev):
        _check(lib.latte_siglip_bwd(_ptr(img_loc), img_loc.stride(0), _ptr(txt_all), txt_all.stride(0),
                                    dt, n_loc, n_all, dim, int(label_offset), _ptr(s), _ptr(b), _ptr(g),
                                    _ptr(d_img), _ptr(d_txt), _DTYPES[gdt], dim, _ptr(d_part), peers,
                                    len(peer_ptrs) if fused else 0, _ptr(d_scale), _ptr(d_bias),
                                    wp, wn, _stream(img_loc)), "latte_siglip_bwd")
    return d_img, (d_part if partial else d_txt), d_scale, d_bias


STAGES = ("fwd_sweep", "fwd_finalize", "bwd_prep", "bwd_sweep", "bwd_gemm", "bwd_finish")


def clip_stage_times(img_loc, txt_loc, img_all, txt_all, label_offset: int, logit_scale,
                     row_lse_all, col_lse_all, reps: int = 5, cross_terms: bool = True,
                     partial: bool = False):
    """Mean milliseconds of every kernel stage of one fwd + bwd (CUDA events on the launching
    stream, recorded inside the library): dict stage name -> ms."""
    lib = load()
    img_loc, txt_loc = _rows(img_loc, "image_features"), _rows(txt_loc, "text_features")
    img_all, txt_all = _rows(img_all, "all_image_features"), _rows(txt_all, "all_text_features")
    dt = _dt(img_loc)
    n_loc, dim = img_loc.shape
    n_all = img_all.shape[0]
    dev = img_loc.device
    s = _scalar_f32(logit_scale)
    g = torch.ones(1, dtype=torch.float32, device=dev)
    row_lse_all = _vec(row_lse_all, torch.float32, "row_lse")
    col_lse_all = _vec(col_lse_all, torch.float32, "col_lse")
    row = torch.empty(n_loc, dtype=torch.float32, device=dev)
    col = torch.empty(n_loc, dtype=torch.float32, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    d_img = torch.empty(n_loc, dim, dtype=img_loc.dtype, device=dev)
    d_txt = torch.empty(n_loc, dim, dtype=img_loc.dtype, device=dev)
    d_part = torch.empty(n_all, dim, dtype=torch.float32, device=dev) if partial else None
    d_scale = torch.empty(1, dtype=torch.float32, device=dev)
    wf = _scratch("fwd", _clip_ws_bytes(n_loc, n_all, dim, dt), dev)
    wb = _scratch("bwd", _clip_ws_bytes(n_loc, n_all, dim, dt, bwd=True), dev)
    wfp, wfn = _aligned_ptr(wf)
    wbp, wbn = _aligned_ptr(wb)
    out = (ctypes.c_float * len(STAGES))()
    with torch.cuda.device(dev):
        _check(lib.latte_clip_stage_times(
            _ptr(img_loc), img_loc.stride(0), _ptr(txt_loc), txt_loc.stride(0), _ptr(img_all),
            img_all.stride(0), _ptr(txt_all), txt_all.stride(0), dt, n_loc, n_all, dim, label_offset,
            _ptr(s), _ptr(row_lse_all), _ptr(col_lse_all), _ptr(row), _ptr(col), _ptr(loss), _ptr(g),
            1.0, int(bool(cross_terms)), _ptr(d_img), _ptr(d_txt), _DTYPES[img_loc.dtype], dim,
            _ptr(d_part), _ptr(d_scale), wfp, wfn, wbp, wbn, _stream(img_loc), int(reps), out),
            "latte_clip_stage_times")
    return {name: float(out[k]) for k, name in enumerate(STAGES)}


# ------------------------------------------------------------------------------ prototypes
def normalize_rows(x: torch.Tensor) -> torch.Tensor:
    x = _rows(x.detach().to(torch.float32), "bank")
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _check(load().latte_normalize_rows(_ptr(x), x.stride(0), _ptr(out), out.stride(0),
                                           x.shape[0], x.shape[1], _stream(x)),
               "latte_normalize_rows")
    return out


def nxc_argmax_margin(x, protos, scale: float = 1.0, row_index=None, want_argmax=True,
                      want_margin=True, want_top1=False):
    """-> (argmax int64[n] | None, margin fp32[n] | None, top1 fp32[n] | None)."""
    x = _rows(x.detach(), "x")
    protos = _rows(protos.detach().to(torch.float32), "prototypes")
    if x.shape[1] != protos.shape[1]:
        raise RuntimeError("nxc: feature dims differ")
    dev = x.device
    if row_index is not None:
        row_index = _vec(row_index, torch.int64, "row_index")
        n = row_index.numel()
    else:
        n = x.shape[0]
    am = torch.empty(n, dtype=torch.int64, device=dev) if want_argmax else None
    mg = torch.empty(n, dtype=torch.float32, device=dev) if want_margin else None
    t1 = torch.empty(n, dtype=torch.float32, device=dev) if want_top1 else None
    wp, wn = _nxc_workspace(_dt(x), row_index is not None, n, x.shape[1], protos.shape[0], dev)
    with torch.cuda.device(dev):
        _check(load().latte_nxc_argmax_margin(_ptr(x), x.stride(0), _dt(x), _ptr(row_index), n,
                                              x.shape[1], _ptr(protos), protos.stride(0),
                                              protos.shape[0], float(scale), _ptr(am), _ptr(mg),
                                              _ptr(t1), wp, wn, _stream(x)),
               "latte_nxc_argmax_margin")
    return am, mg, t1


NXC_MULTI_MAX_CLASSES = 64
NXC_MULTI_MAX_JOBS = 4


class LatteNxcJob(ctypes.Structure):
    """latte_nxc_job_t of include/latte_b200.h."""
    _fields_ = [("x", ctypes.c_void_p), ("ldx", ctypes.c_int64), ("x_dtype", ctypes.c_int),
                ("n", ctypes.c_int64), ("dim", ctypes.c_int64),
                ("planes", ctypes.c_void_p), ("num_classes", ctypes.c_int64),
                ("scale", ctypes.c_float),
                ("argmax_out", ctypes.c_void_p), ("margin_out", ctypes.c_void_p),
                ("top1_out", ctypes.c_void_p)]


class ProtoPlanes:
    """16-bit operand planes of one prototype matrix (latte_nxc_split_prototypes: three bf16 planes for
    bf16 rows, two fp16 planes for fp16 / fp32 rows)."""

    def __init__(self, buf, num_classes, dim, normalized):
        self.buf, self.num_classes, self.dim, self.normalized = buf, num_classes, dim, normalized


def nxc_multi_supported(x: torch.Tensor, num_classes: int) -> bool:
    """The one-launch path takes C <= 64 and rows whose pitch / base are 16-byte aligned."""
    return (num_classes <= NXC_MULTI_MAX_CLASSES and x.dtype in _DTYPES and x.dim() == 2
            and x.stride(1) == 1 and (x.stride(0) * x.element_size()) % 16 == 0
            and x.data_ptr() % 16 == 0 and x.shape[0] > 0)


def nxc_split_prototypes(protos: torch.Tensor, normalize: bool = False,
                         want_normalized: bool = False) -> ProtoPlanes:
    """fp32 [C, D] prototypes -> 16-bit operand planes (once per matrix); ``normalize`` fuses
    F.normalize(protos, dim=1) (train.py:384-389)."""
    protos = _rows(protos.detach().to(torch.float32), "prototypes")
    c, d = protos.shape
    need = ctypes.c_size_t()
    _check(load().latte_nxc_planes_bytes(c, d, ctypes.byref(need)), "latte_nxc_planes_bytes")
    buf = torch.empty(need.value + 128, dtype=torch.uint8, device=protos.device)
    off = (-buf.data_ptr()) % 128
    normed = torch.empty(c, d, dtype=torch.float32, device=protos.device) if want_normalized else None
    with torch.cuda.device(protos.device):
        _check(load().latte_nxc_split_prototypes(_ptr(protos), protos.stride(0), c, d, int(bool(normalize)),
                                                 ctypes.c_void_p(buf.data_ptr() + off), _ptr(normed), d,
                                                 _stream(protos)),
               "latte_nxc_split_prototypes")
    pl = ProtoPlanes(buf, c, d, normed)
    pl.ptr = buf.data_ptr() + off
    return pl


def nxc_multi(jobs):
    """``jobs``: list of dicts {x, planes: ProtoPlanes, scale, argmax, margin, top1 (bools)} sharing
    one operand class (all 16-bit, or all fp32) -> list of (argmax | None, margin | None,
    top1 | None) per job, from ONE launch."""
    if not 1 <= len(jobs) <= NXC_MULTI_MAX_JOBS:
        raise RuntimeError("nxc_multi takes 1..4 jobs")
    arr = (LatteNxcJob * len(jobs))()
    outs = []
    keep = []
    dev = jobs[0]["x"].device
    for k, jb in enumerate(jobs):
        x = jb["x"].detach()
        pl = jb["planes"]
        if x.shape[1] != pl.dim:
            raise RuntimeError("nxc_multi: feature dims differ")
        n = x.shape[0]
        am = torch.empty(n, dtype=torch.int64, device=dev) if jb.get("argmax") else None
        mg = torch.empty(n, dtype=torch.float32, device=dev) if jb.get("margin") else None
        t1 = torch.empty(n, dtype=torch.float32, device=dev) if jb.get("top1") else None
        a = arr[k]
        a.x, a.ldx, a.x_dtype, a.n, a.dim = x.data_ptr(), x.stride(0), _dt(x), n, x.shape[1]
        a.planes, a.num_classes, a.scale = pl.ptr, pl.num_classes, float(jb.get("scale", 1.0))
        a.argmax_out = am.data_ptr() if am is not None else None
        a.margin_out = mg.data_ptr() if mg is not None else None
        a.top1_out = t1.data_ptr() if t1 is not None else None
        outs.append((am, mg, t1))
        keep.append(x)
    with torch.cuda.device(dev):
        _check(load().latte_nxc_multi(arr, len(jobs), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
               "latte_nxc_multi")
    return outs


def _nxc_workspace(dt: int, gathered: bool, n: int, dim: int, classes: int, dev):
    need = ctypes.c_size_t()
    _check(load().latte_nxc_workspace_bytes(dt, int(gathered), n, dim, classes, ctypes.byref(need)),
           "latte_nxc_workspace_bytes")
    if need.value == 0:
        return ctypes.c_void_p(0), 0
    return _aligned_ptr(_scratch("nxc", need.value, dev))


def _seg_workspace(batch: int, dim: int, classes: int, dev):
    need = ctypes.c_size_t()
    _check(load().latte_seg_workspace_bytes(batch, dim, classes, ctypes.byref(need)),
           "latte_seg_workspace_bytes")
    return _aligned_ptr(_scratch("seg", need.value, dev))


def nxc_topk(x, protos, k: int, scale: float = 1.0):
    x = _rows(x.detach(), "x")
    protos = _rows(protos.detach().to(torch.float32), "prototypes")
    n = x.shape[0]
    idx = torch.empty(n, k, dtype=torch.int64, device=x.device)
    val = torch.empty(n, k, dtype=torch.float32, device=x.device)
    wp, wn = _nxc_workspace(_dt(x), False, n, x.shape[1], protos.shape[0], x.device)
    with torch.cuda.device(x.device):
        _check(load().latte_nxc_topk(_ptr(x), x.stride(0), _dt(x), n, x.shape[1], _ptr(protos),
                                     protos.stride(0), protos.shape[0], float(scale), int(k),
                                     _ptr(idx), _ptr(val), wp, wn, _stream(x)),
               "latte_nxc_topk")
    return idx, val


def mix_ema_fwd(class_text, per_image, per_group, bank, preds, zs, w_lbl, w_lbl_zs, w_img, w_grp,
                alpha: float, label_axis: str):
    class_text = _rows(class_text.detach(), "class_text")
    per_image = _rows(per_image.detach(), "per_image")
    per_group = _rows(per_group.detach(), "per_group")
    bank = _rows(bank.detach().to(torch.float32), "bank")
    dt = _dt(class_text)
    if not (_dt(per_image) == _dt(per_group) == dt):
        raise RuntimeError("mix_ema: text sources must share one dtype")
    b, d = per_image.shape
    c = class_text.shape[0]
    dev = per_image.device
    preds, zs = _vec(preds, torch.int64, "preds"), _vec(zs, torch.int64, "zs")
    ws = [_vec(w, torch.float32, "weight") for w in (w_lbl, w_lbl_zs, w_img, w_grp)]
    t_ft = torch.empty(b, d, dtype=per_image.dtype, device=dev)
    t_zs = torch.empty(b, d, dtype=per_image.dtype, device=dev)
    with torch.cuda.device(dev):
        _check(load().latte_mix_ema_fwd(_ptr(class_text), class_text.stride(0), _ptr(per_image),
                                        per_image.stride(0), _ptr(per_group), per_group.stride(0),
                                        _ptr(bank), bank.stride(0), _ptr(preds), _ptr(zs),
                                        _ptr(ws[0]), _ptr(ws[1]), _ptr(ws[2]), _ptr(ws[3]),
                                        float(alpha), LABEL_AXIS[label_axis], dt, b, d, c,
                                        _ptr(t_ft), _ptr(t_zs), d, _stream(per_image)),
               "latte_mix_ema_fwd")
    return t_ft, t_zs


def mix_ema_bwd(d_t_ft, d_t_zs, preds, zs, w_lbl, w_lbl_zs, w_img, w_grp, alpha: float,
                label_axis: str, num_classes: int, want_bank: bool = False):
    """-> (d_class_text fp32 [C, D], d_per_image, d_per_group [B, D], d_bank fp32 [C, D] | None)."""
    d_t_ft = _rows(d_t_ft.detach(), "d_t_ft")
    d_t_zs = _rows(d_t_zs.detach().to(d_t_ft.dtype), "d_t_zs")
    if d_t_zs.stride(0) != d_t_ft.stride(0):
        d_t_ft, d_t_zs = d_t_ft.contiguous(), d_t_zs.contiguous()
    dt = _dt(d_t_ft)
    b, d = d_t_ft.shape
    dev = d_t_ft.device
    preds, zs = _vec(preds, torch.int64, "preds"), _vec(zs, torch.int64, "zs")
    ws = [_vec(w, torch.float32, "weight") for w in (w_lbl, w_lbl_zs, w_img, w_grp)]
    d_ct = torch.zeros(num_classes, d, dtype=torch.float32, device=dev)
    d_pi = torch.empty(b, d, dtype=d_t_ft.dtype, device=dev)
    d_pg = torch.empty(b, d, dtype=d_t_ft.dtype, device=dev)
    d_bank = torch.zeros(num_classes, d, dtype=torch.float32, device=dev) if want_bank else None
    wp, wn = _seg_workspace(b, d, num_classes, dev)
    with torch.cuda.device(dev):
        _check(load().latte_mix_ema_bwd(_ptr(d_t_ft), _ptr(d_t_zs), d_t_ft.stride(0), _ptr(preds),
                                        _ptr(zs), _ptr(ws[0]), _ptr(ws[1]), _ptr(ws[2]), _ptr(ws[3]),
                                        float(alpha), LABEL_AXIS[label_axis], dt, b, d, num_classes,
                                        _ptr(d_ct), d, _ptr(d_pi), _ptr(d_pg), d, _ptr(d_bank), d,
                                        wp, wn, _stream(d_t_ft)),
               "latte_mix_ema_bwd")
    return d_ct, d_pi, d_pg, d_bank


def bank_accumulate(t_ft, t_zs, preds, zs, num_classes: int):
    """-> (sums fp32 [C, D], counts fp32 [C])."""
    t_ft = _rows(t_ft.detach(), "t_ft")
    t_zs = _rows(t_zs.detach().to(t_ft.dtype), "t_zs")
    if t_zs.stride(0) != t_ft.stride(0):
        t_ft, t_zs = t_ft.contiguous(), t_zs.contiguous()
    b, d = t_ft.shape
    dev = t_ft.device
    preds, zs = _vec(preds, torch.int64, "preds"), _vec(zs, torch.int64, "zs")
    sums = torch.empty(num_classes, d, dtype=torch.float32, device=dev)
    counts = torch.empty(num_classes, dtype=torch.float32, device=dev)
    wp, wn = _seg_workspace(b, d, num_classes, dev)
    with torch.cuda.device(dev):
        _check(load().latte_bank_accumulate(_ptr(t_ft), _ptr(t_zs), t_ft.stride(0), _dt(t_ft),
                                            _ptr(preds), _ptr(zs), b, d, num_classes, _ptr(sums), d,
                                            _ptr(counts), wp, wn, _stream(t_ft)),
               "latte_bank_accumulate")
    return sums, counts


def bank_finalize(sums, counts, bank):
    """In-place update of ``bank`` (fp32 [C, D]) for every class with counts > 0."""
    if bank.dtype != torch.float32 or not bank.is_contiguous() or not bank.is_cuda:
        raise RuntimeError("bank_finalize: bank must be a contiguous fp32 CUDA tensor")
    c, d = bank.shape
    with torch.cuda.device(bank.device):
        _check(load().latte_bank_finalize(_ptr(sums), sums.stride(0), _ptr(counts), _ptr(bank),
                                          bank.stride(0), d, c, _stream(bank)),
               "latte_bank_finalize")
    return bank
