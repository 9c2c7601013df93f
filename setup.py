"""Build of the `quant_engine` torch extension for B200 (sm_100a).

Counterpart of the reference's engine/kernels/setup.py:5-25 (setuptools + CUDAExtension("quant_engine")).  The native
code is split in two so that the C-ABI can be bound without torch:
  quantize_b200/libqb200.so      CUDA kernels + C-ABI   (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo)
  quantize_b200/quant_engine.so  pybind11 / ATen shim    (g++, links libqb200.so with an $ORIGIN rpath)

    python setup.py build_ext --inplace      # builds both in-tree (what __graft_entry__.build() does)
    python setup.py install                  # as the reference's README.md:37-42; installs the package + both .so files
"""
import os
import sys

from setuptools import Command, setup
from setuptools.command.build_py import build_py

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


class BuildNative(Command):
    description = "compile libqb200.so (nvcc, sm_100a) and quant_engine.so (g++) in-tree"
    user_options = [("inplace", "i", "ignored: the native build is always in-tree"), ("force", "f", "rebuild")]
    boolean_options = ["inplace", "force"]

    def initialize_options(self):
        self.inplace = 1
        self.force = 0

    def finalize_options(self):
        pass

    def run(self):
        from quantize_b200 import build
        build.build_all(force=bool(self.force))


class BuildPyWithNative(build_py):
    def run(self):
        self.run_command("build_ext")
        super().run()


setup(
    name="quant_engine_b200",
    version="0.1.0",
    description="B200 (sm_100a) implementation of JingInAI/Quantize's quant_engine hot path",
    packages=["quantize_b200"],
    package_data={"quantize_b200": ["libqb200.so", "quant_engine.so", "csrc/*", "../include/qb200.h"]},
    cmdclass={"build_ext": BuildNative, "build_py": BuildPyWithNative},
)
