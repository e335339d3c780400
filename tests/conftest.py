import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "pydrobert-pytorch_b200")
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_cuda = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _small_unit_cost_batches_on_the_wavefront_kernels(monkeypatch):
    """The library sends EVERY unit-cost batch with references of at most 64 tokens to the
    short-reference bit-vector kernel (B200LEV_BVSHORT_MIN_PAIRS defaults to 1).  Most tests here
    use tiny unit-cost batches to exercise the pack / wavefront / group / CTA kernels, so the suite
    runs with the threshold at 64; the `bv_form` tests (and the reference's own suite, which sets
    nothing) cover the default."""
    if "B200LEV_BVSHORT_MIN_PAIRS" not in os.environ:
        monkeypatch.setenv("B200LEV_BVSHORT_MIN_PAIRS", "64")


class Golden:
    """Accessor for one tests/golden/*.npz fixture (written by make_golden.py)."""

    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name))
        self.params = json.loads(str(self.z["params"])) if "params" in self.z.files else {}

    def names(self, prefix=""):
        return [k for k in self.params if k.startswith(prefix)]

    def has(self, case, field):
        return f"{case}.{field}" in self.z.files

    def get(self, case, field):
        return self.z[f"{case}.{field}"]


@pytest.fixture(scope="session")
def golden_sm():
    return Golden("string_matching.npz")


@pytest.fixture(scope="session")
def golden_loss():
    return Golden("losses.npz")


@pytest.fixture(scope="session")
def golden_seqlp():
    return Golden("seqlp.npz")


@pytest.fixture(scope="session")
def golden_ctc():
    return Golden("ctc.npz")


@pytest.fixture(scope="session")
def golden_seqlp_packed():
    return Golden("seqlp_packed.npz")


@pytest.fixture(scope="session")
def golden_decode():
    return Golden("decode.npz")


@pytest.fixture(scope="session")
def golden_scoring():
    with open(os.path.join(GOLDEN, "scoring.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_sclite():
    return np.load(os.path.join(GOLDEN, "sclite.npz"))
