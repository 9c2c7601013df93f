"""quantize_b200 — B200 (sm_100a) implementation of the quantized-operator hot path of JingInAI/Quantize.

Layout
  csrc/          CUDA kernels + the C-ABI (include/qb200.h) + the `quant_engine` torch-extension shim
  build.py       in-tree build (nvcc / g++), used by __graft_entry__.build() and setup.py
  capi.py        ctypes binding of libqb200.so (what a non-torch host would bind; used by the parity tests)
  engine.py      loader of the `quant_engine` extension — the module the reference's engine/__init__.py imports
  host.py        host-side mirror of the reference's quantized conv layer (calibrate / pack / forward through the op)
  models.py      synthetic-weight CNN stacks of BASELINE.json's configs built from host.py layers
  dist.py        batch-sharded multi-GPU driver (one process per GPU; NCCL only to gather results)

There is no CPU fallback anywhere in this package: importing works without a GPU (so the build can be checked),
every compute call raises without one.
"""
__version__ = "0.1.0"
