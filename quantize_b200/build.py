"""In-tree build of the native code (sm_100a only).

Produces, next to this file:
  libqb200.so       the C-ABI library (include/qb200.h): nvcc, no torch dependency
  quant_engine*.so  the Python-facing torch extension with the reference's 8 op names
                    (reference engine/kernels/pybind.cpp:7-17, engine/kernels/setup.py:5-25): g++ only,
                    links against libqb200.so with an $ORIGIN rpath

The built files are git-ignored but travel to the GPU box with the repo snapshot.
"""
import hashlib
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libqb200.so")
EXT = os.path.join(HERE, "quant_engine.so")

CU_SOURCES = ["common.cu", "pack.cu", "actquant.cu", "wprep.cu", "conv_direct.cu", "conv_umma.cu", "conv_dw.cu", "conv_api.cu", "pool.cu", "packed_ops.cu", "calib.cu"]
HEADERS = ["common.cuh", "conv_common.cuh", "quant_math.cuh", os.path.join(ROOT, "include", "qb200.h")]

NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout[-3000:], r.stderr[-6000:]))
    return r.stdout + r.stderr


def _digest(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale(target, stamp_name, digest):
    stamp = os.path.join(BUILD, stamp_name)
    if not os.path.exists(target) or not os.path.exists(stamp):
        return True
    with open(stamp) as f:
        return f.read().strip() != digest


def _write_stamp(stamp_name, digest):
    with open(os.path.join(BUILD, stamp_name), "w") as f:
        f.write(digest)


def build_lib(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES]
    digest = _digest(srcs + hdrs, " ".join(NVCC_FLAGS))
    if not force and not _stale(LIB, "lib.stamp", digest):
        return LIB
    objs = [os.path.join(BUILD, s.replace(".cu", ".o")) for s in CU_SOURCES]

    def compile_one(pair):
        src, obj = pair
        return _run(["nvcc"] + NVCC_FLAGS + ["-c", src, "-o", obj])

    with ThreadPoolExecutor(len(srcs)) as ex:
        logs = list(ex.map(compile_one, zip(srcs, objs)))
    with open(os.path.join(BUILD, "ptxas.log"), "w") as f:
        f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    _run(["nvcc", "-shared", "-o", LIB] + objs + ["-lcudart"])
    _write_stamp("lib.stamp", digest)
    return LIB


def build_ext(force=False):
    """quant_engine torch extension (pybind11).  Only host C++ — compiled with g++."""
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(CSRC, "pybind.cpp")
    digest = _digest([src, os.path.join(ROOT, "include", "qb200.h")], torch.__version__ + " rpath=$ORIGIN:$ORIGIN/quantize_b200")
    if not force and not _stale(EXT, "ext.stamp", digest):
        return EXT
    inc = [f"-I{p}" for p in ce.include_paths()] + [f"-I{sysconfig.get_paths()['include']}",
                                                   "-I/usr/local/cuda/include", f"-I{os.path.join(ROOT, 'include')}"]
    defs = ["-DTORCH_EXTENSION_NAME=quant_engine", "-DTORCH_API_INCLUDE_EXTENSION_H",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    obj = os.path.join(BUILD, "pybind.o")
    _run(["g++", "-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden"] + inc + defs + ["-c", src, "-o", obj])
    _run(["g++", "-shared", "-o", EXT, obj, f"-L{HERE}", "-lqb200", f"-L{libdir}", "-L/usr/local/cuda/lib64",
          "-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart",
          "-Wl,-rpath,$ORIGIN:$ORIGIN/quantize_b200", f"-Wl,-rpath,{libdir}"])
    _write_stamp("ext.stamp", digest)
    return EXT


def top_level_name():
    """file name of the extension as a top-level module: quant_engine.<abi tag>.so"""
    return "quant_engine" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so")


def install_top_level(dst=None):
    """Copy quant_engine.so to `dst` (default: the repository root) as the TOP-LEVEL module `quant_engine`, the name the
    reference imports (engine/__init__.py:1-5: `from quant_engine import *`).  Its rpath also lists
    $ORIGIN/quantize_b200, so it finds libqb200.so inside the package directory beside it."""
    import shutil
    dst = dst or os.path.join(ROOT, top_level_name())
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(EXT) or os.path.getsize(dst) != os.path.getsize(EXT):
        shutil.copy2(EXT, dst)
    return dst


def build_all(force=False, verbose=False):
    build_lib(force=force, verbose=verbose)
    build_ext(force=force)
    install_top_level()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
    print(EXT)
