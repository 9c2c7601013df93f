"""CPU: the C-ABI library and the torch extension build, load and export what include/qb200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from quantize_b200 import build
    build.build_all()
    return build


def _declared_functions():
    with open(os.path.join(ROOT, "include", "qb200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qb200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built):
    from quantize_b200 import capi
    L = capi.lib()
    declared = _declared_functions()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/qb200.h but not exported by libqb200.so"
    assert sorted(capi.SYMBOLS) == declared
    assert L.qb200_version() >= 100


def test_pure_host_entry_points(built):
    from quantize_b200 import capi
    L = capi.lib()
    assert L.qb200_packed_bytes(10, 3) == 4          # tpack.cu:224 ceil(numel*n/8)
    assert L.qb200_packed_bytes(8, 8) == 8
    assert L.qb200_packed_bytes(1, 9) == -1
    assert L.qb200_padded_channels(3) == 32 and L.qb200_padded_channels(64) == 64 and L.qb200_padded_channels(65) == 96
    s = capi.conv_shape(2, 3, 224, 224, 64, 3, 7, 7, 2, 3, 8, 1)
    assert capi.conv_out_hw(s) == (112, 112)         # quantconv2d_float_input.cu:178-179
    s = capi.conv_shape(1, 64, 56, 56, 64, 64, 3, 3, 1, 1, 4, 1)
    assert capi.conv_out_hw(s) == (56, 56)
    assert 56 * 56 * 64 <= L.qb200_conv_workspace_bytes(s) <= 58 * 58 * 64 + 256   # zero-padded NHWC for the halo path
    assert L.qb200_conv_prepared_bytes(s) >= 64 * 9 * 64 + 64 * 16 * 4
    bad = capi.conv_shape(1, 64, 56, 56, 64, 48, 3, 3, 1, 1, 4, 1)
    with pytest.raises(capi.Qb200Error, match="not divisible"):
        capi.conv_out_hw(bad)
    bad = capi.conv_shape(1, 64, 56, 56, 64, 64, 3, 3, 1, 1, 9, 1)
    with pytest.raises(capi.Qb200Error, match=r"\(0, 8\]"):
        capi.conv_out_hw(bad)


def test_extension_exports_reference_names(built, engine):
    # engine/kernels/pybind.cpp:9-16 — all 8 names, or `import modelzoo` of the reference breaks
    for name in ["tpack", "tunpack", "linear", "quantlinear", "quantlinear_float_input", "conv2d", "quantconv2d",
                 "quantconv2d_float_input"]:
        assert callable(getattr(engine, name))
    assert engine._abi_version() >= 100


def test_no_cpu_fallback(engine):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        engine.tpack(torch.zeros(8), 4, True)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):   # same message as the reference's CHECK_CUDA
        engine.quantconv2d_float_input(torch.zeros(1, 1, 3, 3), torch.zeros(1, dtype=torch.uint8),
                                       torch.zeros(6, dtype=torch.int32), torch.ones(1), torch.zeros(1), None, 1, 0)
    # argument errors are the reference's (tpack.cu:13)
    with pytest.raises(RuntimeError, match=r"n_bits must be in the range \(0, 8\]"):
        engine.tpack(torch.zeros(8), 0, True)
    with pytest.raises(RuntimeError, match="outside the quantized-operator path"):
        engine.linear(torch.zeros(2, 2), torch.zeros(2, 2))
    # the packed-activation ops check their tensors like the reference (quantconv2d.cu:178-189, quantlinear.cu:243-250)
    u8, i32, f1 = torch.zeros(4, dtype=torch.uint8), torch.zeros(6, dtype=torch.int32), torch.ones(1)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        engine.quantconv2d(u8, i32, f1, f1, u8, i32, f1, f1, None, 1, 0)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        engine.quantlinear(u8, i32[:4], f1, f1, u8, i32[:4], f1, f1, None)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        engine.minmax(torch.zeros(4), 0, 0, False)


def test_setup_py_yields_top_level_quant_engine(built, tmp_path):
    """reference engine/kernels/setup.py:5-25 + README.md:37-42: after the build, `import quant_engine` works from a clean
    interpreter (the reference's engine/__init__.py:1-5 does `from quant_engine import *` after `import torch`)."""
    import subprocess
    import sys
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    r = subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=ROOT, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    code = ("import torch, quant_engine, os; assert quant_engine._abi_version() >= 200; "
            "names = ['tpack','tunpack','linear','quantlinear','quantlinear_float_input','conv2d','quantconv2d','quantconv2d_float_input']; "
            "assert all(callable(getattr(quant_engine, n)) for n in names); print(os.path.basename(quant_engine.__file__))")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().startswith("quant_engine.")
    # `setup.py build` (what `install` runs first): the module lands at the top level of the build tree, beside the package
    base = str(tmp_path / "b")
    r = subprocess.run([sys.executable, "setup.py", "build", "--build-base", base], cwd=ROOT, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    libdirs = [d for d in os.listdir(base) if d.startswith("lib")]
    assert libdirs, os.listdir(base)
    lib = os.path.join(base, libdirs[0])
    assert any(f.startswith("quant_engine.") and f.endswith(".so") for f in os.listdir(lib)), os.listdir(lib)
    assert os.path.exists(os.path.join(lib, "quantize_b200", "libqb200.so"))
    env2 = dict(env, PYTHONPATH=lib)
    r = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), env=env2, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under quantize_b200/ may reference it."""
    pkg = os.path.join(ROOT, "quantize_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath:
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    text = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert "qoracle" not in text, fn
