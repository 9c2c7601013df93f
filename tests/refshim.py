"""TEST INFRASTRUCTURE: import the UNMODIFIED reference Python package from /root/reference.

Only usable in the build container (the GPU box has no /root/reference): used by tests/golden/make_golden.py to
generate the committed fixtures and by the container-only cross-checks (skipped elsewhere).
The reference needs `robustbench` and `ftfy` (requirements.txt:1-3; absent here) only for model families that are
off the hot path, so they are stubbed; `quant_engine` is whatever engine module the caller injects.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("QB200_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "modelzoo"))


def load_reference(engine_module):
    """Returns the reference's `modelzoo.modules` package with `engine_module` serving as `quant_engine`."""
    if not available():
        raise FileNotFoundError(REF_ROOT)
    if "robustbench" not in sys.modules:
        rb = types.ModuleType("robustbench")
        rb.load_model = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("robustbench stub"))
        sys.modules["robustbench"] = rb
    if "ftfy" not in sys.modules:
        ft = types.ModuleType("ftfy")
        ft.fix_text = lambda s: s
        sys.modules["ftfy"] = ft
    # the reference needs all 8 names importable; fill what the given engine lacks with stubs
    proxy = types.ModuleType("quant_engine")
    names = ["tpack", "tunpack", "linear", "quantlinear", "quantlinear_float_input", "conv2d", "quantconv2d",
             "quantconv2d_float_input"]

    def _missing(name):
        def f(*a, **k):
            raise RuntimeError(f"quant_engine.{name} is not provided by the injected engine")
        return f

    for n in names:
        setattr(proxy, n, getattr(engine_module, n, None) or _missing(n))
    proxy.__all__ = names
    sys.modules["quant_engine"] = proxy
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    for name in [n for n in sys.modules if n == "engine" or n.startswith("engine.")]:
        del sys.modules[name]
    import modelzoo.modules as mods  # noqa: E402
    return mods


def oracle_engine_module():
    """A `quant_engine` stand-in for the CPU container: tpack/tunpack from the oracle's C restatement, the
    remaining six names present but unusable (the reference only needs them importable)."""
    import numpy as np
    import torch
    import oracle

    m = types.ModuleType("quant_engine")

    def tpack(x, n_bits, sign):
        packed, des = oracle.tpack(x.detach().cpu().numpy(), n_bits, sign)
        return [torch.from_numpy(packed).to(x.device), torch.from_numpy(des).to(x.device)]

    def tunpack(x, des):
        out = oracle.tunpack(x.detach().cpu().numpy(), des.detach().cpu().numpy())
        return torch.from_numpy(np.ascontiguousarray(out)).to(x.device)

    def _nope(*a, **k):
        raise RuntimeError("not available in the oracle engine stand-in")

    m.tpack, m.tunpack = tpack, tunpack
    for n in ("linear", "quantlinear", "quantlinear_float_input", "conv2d", "quantconv2d", "quantconv2d_float_input"):
        setattr(m, n, _nope)
    m.__all__ = ["tpack", "tunpack", "linear", "quantlinear", "quantlinear_float_input", "conv2d", "quantconv2d",
                 "quantconv2d_float_input"]
    return m
