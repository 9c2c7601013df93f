"""Build of the `quant_engine` torch extension for B200 (sm_100a).

Counterpart of the reference's engine/kernels/setup.py:5-25 (setuptools + CUDAExtension("quant_engine") + BuildExtension):
the same commands produce a TOP-LEVEL module `quant_engine`, which the reference's engine/__init__.py:1-5 imports.

    python setup.py build_ext --inplace      # ./quant_engine.<abi>.so + quantize_b200/{libqb200,quant_engine}.so
    python setup.py install                  # as the reference's README.md:37-42: quant_engine + the quantize_b200 package

The native code is split in two so that the C-ABI can be bound without torch:
  quantize_b200/libqb200.so   CUDA kernels + C-ABI (include/qb200.h): nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo
  quant_engine.<abi>.so       pybind11 / ATen shim: g++, links libqb200.so; rpath $ORIGIN:$ORIGIN/quantize_b200
torch.utils.cpp_extension.CUDAExtension would compile every .cu with torch headers and torch's arch list; the explicit
nvcc recipe (quantize_b200/build.py) keeps the kernels torch-free and pins sm_100a.
"""
import os
import sys

from setuptools import Extension, setup
from setuptools.command.build_ext import build_ext

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


class BuildNative(build_ext):
    """builds both shared objects in-tree (stamped: a no-op when sources are unchanged), then places the extension where
    setuptools expects the module `quant_engine` (the build directory, or the source root with --inplace)."""

    def build_extension(self, ext):
        from quantize_b200 import build
        build.build_lib(force=bool(self.force))
        build.build_ext(force=bool(self.force))
        build.install_top_level(self.get_ext_fullpath(ext.name))


setup(
    name="quant_engine_b200",
    version="0.2.0",
    description="B200 (sm_100a) implementation of JingInAI/Quantize's quant_engine hot path",
    packages=["quantize_b200"],
    package_data={"quantize_b200": ["libqb200.so", "quant_engine.so", "csrc/*"]},
    ext_modules=[Extension("quant_engine", sources=[])],
    cmdclass={"build_ext": BuildNative},
)
