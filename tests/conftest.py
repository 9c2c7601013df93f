import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_pack_kat():
    with open(os.path.join(GOLDEN, "pack_kat.json")) as f:
        return json.load(f)["cases"]


def conv_fixture_names():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "conv_*.npz")))


def load_conv_fixture(name):
    d = dict(np.load(os.path.join(GOLDEN, f"conv_{name}.npz")))
    d["stride"], d["pad"], d["groups"] = int(d["stride"][0]), int(d["pad"][0]), int(d["groups"][0])
    d["bias"] = d["bias"] if d["bias"].size else None
    return d


@pytest.fixture(scope="session")
def engine():
    """The product's torch extension; building is part of the contract (__graft_entry__.build)."""
    import torch  # noqa: F401
    from quantize_b200 import engine as eng
    return eng.load()
