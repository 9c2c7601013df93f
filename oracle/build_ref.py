"""TEST INFRASTRUCTURE ONLY: compile the UNMODIFIED reference hot-path sources into oracle/_ref/.

Recipe (no reference build system is run; the sources are compiled where they lie):
  nvcc  /root/reference/engine/kernels/tpack/tpack.cu
  nvcc  /root/reference/engine/kernels/functions/quantconv2d_float_input.cu
  nvcc  /root/reference/engine/kernels/functions/quantlinear_float_input.cu
  nvcc  /root/reference/engine/kernels/functions/quantconv2d.cu
  nvcc  /root/reference/engine/kernels/functions/quantlinear.cu
  g++   oracle/ref_bind.cpp   (our own 6-op pybind shim, includes the reference headers)
  link  -> oracle/_ref/quant_engine_ref.so   (git-ignored; travels to the GPU box with gpurun)

The reference tpack/tunpack have a CPU path (tpack.cu:140-190, :371-419) so they run in the CPU
container; quantconv2d_float_input is CUDA-only (quantconv2d_float_input.cu:151) and is only usable
on the GPU box, where it is the op-level oracle for the `-m gpu` tests.
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("QB200_REFERENCE_ROOT", "/root/reference")
KERN = os.path.join(REF, "engine", "kernels")
OUT = os.path.join(HERE, "_ref")
NAME = "quant_engine_ref"


def _torch_flags():
    import torch
    from torch.utils import cpp_extension as ce
    inc = [f"-I{p}" for p in ce.include_paths()] + [f"-I{sysconfig.get_paths()['include']}", f"-I{KERN}",
                                                   "-I/usr/local/cuda/include"]
    defs = [f"-DTORCH_EXTENSION_NAME={NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    return inc, defs, libdir


def available():
    return os.path.exists(os.path.join(OUT, NAME + ".so"))


def build(force=False):
    so = os.path.join(OUT, NAME + ".so")
    shim = os.path.join(HERE, "ref_bind.cpp")
    stale = os.path.exists(so) and os.path.isdir(KERN) and max(os.path.getmtime(shim), os.path.getmtime(__file__)) > os.path.getmtime(so)
    if os.path.exists(so) and not force and not stale:
        return so
    if not os.path.isdir(KERN):
        raise FileNotFoundError(f"reference sources not present at {KERN}; oracle/_ref must be prebuilt")
    os.makedirs(OUT, exist_ok=True)
    inc, defs, libdir = _torch_flags()
    nvcc = ["nvcc", "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
            "-gencode", "arch=compute_100,code=[sm_100,compute_100]",
            "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
            "-D__CUDA_NO_BFLOAT16_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__"] + inc + defs
    gxx = ["g++", "-O2", "-std=c++17", "-fPIC"] + inc + defs
    funcs = ["quantconv2d_float_input", "quantlinear_float_input", "quantconv2d", "quantlinear"]
    jobs = [(nvcc + ["-c", os.path.join(KERN, "tpack", "tpack.cu"), "-o", os.path.join(OUT, "tpack.o")])]
    jobs += [(nvcc + ["-c", os.path.join(KERN, "functions", f + ".cu"), "-o", os.path.join(OUT, f + ".o")]) for f in funcs]
    jobs += [(gxx + ["-c", os.path.join(HERE, "ref_bind.cpp"), "-o", os.path.join(OUT, "ref_bind.o")])]
    with ThreadPoolExecutor(6) as ex:
        for r in ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs):
            if r.returncode != 0:
                raise RuntimeError("reference build failed:\n" + r.stderr[-4000:])
    objs = ["tpack.o"] + [f + ".o" for f in funcs] + ["ref_bind.o"]
    link = ["g++", "-shared", "-o", so] + [os.path.join(OUT, o) for o in objs] + [
            f"-L{libdir}", "-L/usr/local/cuda/lib64", "-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python",
            "-lc10_cuda", "-ltorch_cuda", "-lcudart", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference link failed:\n" + r.stderr[-4000:])
    for o in objs:
        os.remove(os.path.join(OUT, o))
    return so


def load():
    """Import oracle/_ref/quant_engine_ref.so (torch must be imported first for libtorch symbols)."""
    import importlib.util
    import torch  # noqa: F401
    so = os.path.join(OUT, NAME + ".so")
    if not os.path.exists(so):
        raise FileNotFoundError(so)
    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
