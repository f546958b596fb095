"""CPU-only checks of the C-ABI boundary: the shared library builds, loads without a GPU
and exports exactly the symbols include/latte_b200.h declares."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from latteclip_b200 import _lib
    _lib.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "latte_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(latte_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 13
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in latte_b200.h but not exported"


def test_python_binding_covers_header():
    from latteclip_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_symbols()


def test_version_and_status_strings(lib):
    assert b"sm_100a" in lib.latte_version()
    assert lib.latte_status_string(0) == b"ok"
    for code in (-1, -2, -3, -4, -5):
        assert len(lib.latte_status_string(code)) > 3


def test_bad_arguments_are_rejected_without_a_gpu(lib):
    n = ctypes.c_size_t(0)
    assert lib.latte_clip_workspace_bytes(0, 0, 512, 1, ctypes.byref(n)) == -1
    assert lib.latte_clip_workspace_bytes(256, 256, 512, 7, ctypes.byref(n)) == -1
    assert lib.latte_clip_workspace_bytes(256, 1024, 512, 1, ctypes.byref(n)) == 0
    assert n.value > 0
    # null pointers never reach a kernel launch
    assert lib.latte_normalize_rows(None, 512, None, 512, 4, 512, None) == -1
    assert lib.latte_clip_fwd(None, 0, None, 0, None, 0, None, 0, 1, 1, 1, 1, 0,
                              None, None, None, None, None, None, None, None, 0, None) == -1
    assert lib.latte_clip_fwd_rows(None, 0, None, 0, 1, 1, 1, 8, 0, None, None, None, None, None,
                                   None, 0, None) == -1
    assert lib.latte_clip_rank_sweep_supported(1, 512) == 1      # bf16, dim 512
    assert lib.latte_clip_rank_sweep_supported(0, 512) == 0      # fp32 features: two-sweep path
    assert lib.latte_clip_rank_sweep_supported(1, 768) == 1      # ViT-L/14 width
    assert lib.latte_clip_rank_sweep_supported(1, 1024) == 0


def test_product_path_has_no_cpu_fallback():
    import torch
    import latteclip_b200 as lb
    loss = lb.ClipLoss()
    x = torch.randn(8, 16)
    with pytest.raises(RuntimeError, match="CUDA"):
        loss(x, x, torch.tensor(10.0))


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "latteclip_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_siglip_entry_points_reject_bad_arguments_without_a_gpu(lib):
    n = ctypes.c_size_t(0)
    assert lib.latte_siglip_supported(1, 512) == 1 and lib.latte_siglip_supported(2, 768) == 1
    assert lib.latte_siglip_supported(0, 512) == 0          # fp32 features
    assert lib.latte_siglip_supported(1, 60) == 0 and lib.latte_siglip_supported(1, 1024) == 0
    assert lib.latte_siglip_workspace_bytes(256, 1024, 512, 1, 1, 0, ctypes.byref(n)) == 0 and n.value > 0
    fwd_only = ctypes.c_size_t(0)
    assert lib.latte_siglip_workspace_bytes(256, 1024, 512, 1, 0, 0, ctypes.byref(fwd_only)) == 0
    assert fwd_only.value < n.value
    assert lib.latte_siglip_workspace_bytes(256, 128, 512, 1, 0, 0, ctypes.byref(n)) == -1    # n_all < n_loc
    assert lib.latte_siglip_workspace_bytes(256, 256, 512, 0, 0, 0, ctypes.byref(n)) != 0     # fp32
    assert lib.latte_siglip_fwd(None, 0, None, 0, 1, 1, 1, 8, 0, None, None, None, None, 0, None) == -1
    assert lib.latte_siglip_bwd(None, 0, None, 0, 1, 1, 1, 8, 0, None, None, None, None, None, 1, 8,
                                None, None, 0, None, None, None, 0, None) == -1


def test_factory_and_module_contracts_on_cpu():
    """create_loss mirrors factory.py:323-351; argument errors surface before any kernel launch."""
    from types import SimpleNamespace
    import torch
    import latteclip_b200 as lb
    base = dict(model="ViT-B-32", siglip=False, distill=False, local_loss=True, gather_with_grad=True,
                rank=0, world_size=1, horovod=False)
    loss = lb.create_loss(SimpleNamespace(**base))
    assert isinstance(loss, lb.ClipLoss) and loss.local_loss and loss.gather_with_grad and loss.cache_labels
    assert len(list(loss.parameters())) == 0
    sig = lb.create_loss(SimpleNamespace(**dict(base, siglip=True, rank=3, world_size=8)))
    assert isinstance(sig, lb.SigLipLoss) and sig.rank == 3 and sig.world_size == 8
    assert len(list(sig.parameters())) == 0
    with pytest.raises(NotImplementedError):
        lb.create_loss(SimpleNamespace(**dict(base, model="coca_ViT-B-32")))
    # the reference factory has no distillation branch: args.distill still yields a ClipLoss
    assert type(lb.create_loss(SimpleNamespace(**dict(base, distill=True)))) is lb.ClipLoss
    assert issubclass(lb.DistillClipLoss, lb.ClipLoss)
    with pytest.raises(NotImplementedError):
        lb.DistillClipLoss(rank=0, world_size=2)(x16 := torch.zeros(8, 16), x16, 1.0, x16, x16, 1.0)
    with pytest.raises(AssertionError):
        lb.create_loss(SimpleNamespace(**dict(base, siglip=True, horovod=True)))
    with pytest.raises(AssertionError):
        lb.SigLipLoss(use_horovod=True)
    x = torch.nn.functional.normalize(torch.randn(16, 64), dim=1)
    with pytest.raises(RuntimeError):                       # fp32 features: no silent cast, no fallback
        sig(x, x, torch.tensor(10.0), torch.tensor(-10.0))
    with pytest.raises(RuntimeError):                       # shape mismatch
        sig(x.bfloat16(), x.bfloat16()[:8], torch.tensor(10.0), torch.tensor(-10.0))
    with pytest.raises(RuntimeError):
        loss(x, x[:8], torch.tensor(100.0))
    # materialising utilities stay plain torch
    z = sig.get_logits(x, x, torch.tensor(10.0), torch.tensor(-10.0))
    lab = sig.get_ground_truth(z.device, z.dtype, 16)
    assert z.shape == (16, 16) and float(lab.diagonal().min()) == 1.0 and float(lab.sum()) == 16 - 240


def test_setup_py_build_hook_produces_the_library(tmp_path):
    """`python setup.py build_py` (the build hook north_star asks for; the reference's
    setup.py:22-61 has no native step) compiles csrc/ with nvcc for sm_100a and ships the C-ABI
    library inside the package."""
    import subprocess
    import sys
    out = tmp_path / "lib"
    res = subprocess.run([sys.executable, "setup.py", "-q", "build_py", "--build-lib", str(out)], cwd=ROOT,
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout + res.stderr
    so = out / "latteclip_b200" / "_C" / "liblatte_b200.so"
    assert so.exists()
    lib = ctypes.CDLL(str(so))
    lib.latte_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.latte_version()


def test_workspace_size_queries_allocate_nothing(lib):
    """Every scratch buffer is caller-provided (include/latte_b200.h conventions): the size
    functions answer without a GPU, and the sources never allocate device memory themselves."""
    n = ctypes.c_size_t(0)
    assert lib.latte_nxc_workspace_bytes(0, 0, 32768, 512, 47, ctypes.byref(n)) == 0 and n.value > 0
    assert lib.latte_nxc_workspace_bytes(0, 0, 64, 512, 47, ctypes.byref(n)) == 0 and n.value == 0
    assert lib.latte_seg_workspace_bytes(32768, 512, 47, ctypes.byref(n)) == 0 and n.value > 0
    assert lib.latte_clip_fwd_rank_workspace_bytes(4096, 32768, 512, 2, ctypes.byref(n)) == 0 and n.value > 0
    assert lib.latte_clip_bwd_workspace_bytes(4096, 32768, 512, 2, ctypes.byref(n)) == 0
    assert n.value > 4096 * 32768 * 2          # holds the fp16 gradient weights G [n_loc, N]
    csrc = os.path.join(ROOT, "latteclip_b200", "csrc")
    for f in os.listdir(csrc):
        src = open(os.path.join(csrc, f)).read()
        for banned in ("cudaMalloc", "cudaFree", "cudaMemPool"):
            assert banned not in src, f"{f} calls {banned}*: scratch must come from the caller"
