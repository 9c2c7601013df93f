"""Loader of the `quant_engine` torch extension (the module the reference's engine/__init__.py:1-5 imports).

`load()` returns the extension module and registers it as `sys.modules['quant_engine']`, so the reference's
`from quant_engine import *` resolves to this engine.  A missing build raises — there is no Python fallback
(the reference's own fallback is dead code).
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
EXT_PATH = os.path.join(HERE, "quant_engine.so")
OPS = ["tpack", "tunpack", "linear", "quantlinear", "quantlinear_float_input", "conv2d", "quantconv2d",
       "quantconv2d_float_input"]

_mod = None


def load():
    global _mod
    if _mod is None:
        import torch  # noqa: F401  (libtorch symbols must be loaded first, as in the reference)
        top = sys.modules.get("quant_engine")
        if top is not None and hasattr(top, "_abi_version"):   # already imported as the top-level module (setup.py)
            _mod = top
            return _mod
        if not os.path.exists(EXT_PATH) or not os.path.exists(os.path.join(HERE, "libqb200.so")):
            raise ImportError("quant_engine is not built: run `python -m quantize_b200.build` "
                              "(or __graft_entry__.build()); this engine has no fallback implementation")
        spec = importlib.util.spec_from_file_location("quant_engine", EXT_PATH)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.__all__ = list(OPS)
        sys.modules.setdefault("quant_engine", mod)
        _mod = mod
    return _mod
