"""Packaging of latteclip_b200 with the CUDA build hook BASELINE.json's north_star asks for.

The reference's setup.py (/root/reference/setup.py:22-61) has no ext_modules: it is pure Python.
This one adds ONE build step -- nvcc compiles latteclip_b200/csrc/*.cu for sm_100a into the C-ABI
shared library latteclip_b200/_C/liblatte_b200.so -- hooked into build_py so that `pip install .`,
`python setup.py build_py` and `python setup.py develop` all produce it.  The library is loaded with
ctypes (latteclip_b200/_lib.py); there is no torch C++ extension and no libtorch linkage.
"""

import importlib.util
import os

from setuptools import find_packages, setup
from setuptools.command.build_py import build_py

HERE = os.path.dirname(os.path.abspath(__file__))


def _load_build_recipe():
    # by path: importing the package would import torch
    spec = importlib.util.spec_from_file_location(
        "_latte_build", os.path.join(HERE, "latteclip_b200", "_build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class BuildPyWithCuda(build_py):
    """build_py + nvcc -gencode arch=compute_100a,code=sm_100a (see latteclip_b200/_build.py)."""

    def run(self):
        path = _load_build_recipe().build(force=bool(os.environ.get("LATTE_B200_FORCE_BUILD")))
        self.announce(f"built {path}", level=2)
        super().run()


setup(
    name="latteclip_b200",
    version="0.2.0",
    description="B200-native (sm_100a) loss head for LatteCLIP: ClipLoss / SigLipLoss / prototype path "
                "as hand-written CUDA behind a C ABI",
    packages=find_packages(include=["latteclip_b200", "latteclip_b200.*"]),
    package_data={"latteclip_b200": ["_C/*.so", "csrc/*.cu", "csrc/*.cuh"]},
    data_files=[("include", ["include/latte_b200.h"])],
    python_requires=">=3.9",
    install_requires=["torch>=2.4"],
    cmdclass={"build_py": BuildPyWithCuda},
    zip_safe=False,
)
