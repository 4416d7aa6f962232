"""pytest configuration: markers, import paths, shared fixtures."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gstools-core_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def kat():
    """The reference's known-answer vectors (src/field.rs:258-431)."""
    with open(os.path.join(ROOT, "tests", "golden", "field_rs_kat.json")) as f:
        raw = json.load(f)
    return {k: np.array(v, dtype=np.float64) for k, v in raw.items() if not k.startswith("_")}


def ulp_diff(a, b):
    """Distance in units in the last place between two float64 arrays (same sign assumed)."""
    a = np.ascontiguousarray(a, dtype=np.float64).view(np.int64)
    b = np.ascontiguousarray(b, dtype=np.float64).view(np.int64)
    return np.abs(a - b)
