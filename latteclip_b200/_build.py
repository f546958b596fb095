"""nvcc build recipe of liblatte_b200.so (no torch import: setup.py loads this file by path).

    python -m latteclip_b200.build          # or: pip install . / python setup.py build_py
"""

import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_DIR = os.path.join(_HERE, "_C")
SO_PATH = os.path.join(SO_DIR, "liblatte_b200.so")
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["api.cu", "clip_tc.cu", "clip_pair.cu", "clip_simt.cu", "nxc.cu", "nxc_tc.cu", "nxc_stream.cu", "proto.cu", "siglip.cu"]
HEADERS = ["latte_common.cuh", "tc_ptx.cuh"]
PUBLIC_HEADER = os.path.join(os.path.dirname(_HERE), "include", "latte_b200.h")
# sm_100a only: tcgen05 / TMEM / TMA code paths have no other target
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into latteclip_b200/_C/liblatte_b200.so (in-tree, so the
    built library travels with the repository snapshot).  Rebuilds only when a source is newer."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS] + [PUBLIC_HEADER]
    if not force and os.path.exists(SO_PATH):
        so_m = os.path.getmtime(SO_PATH)
        if all(os.path.getmtime(d) <= so_m for d in deps):
            return SO_PATH
    os.makedirs(SO_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", SO_PATH] + srcs
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return SO_PATH
