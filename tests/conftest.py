import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def load_golden(name):
    """tests/golden/<name>.npz -> {case: {key: array}} (keys are 'case/key')."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {}
    for k in z.files:
        case, key = k.split("/", 1)
        out.setdefault(case, {})[key] = z[k]
    return out


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    orc.build()
    return orc


def bits_equal(a, b):
    """fp32 bit equality (treating +0/-0 as equal, NaN==NaN)."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def n_mismatch(a, b):
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    return int(np.sum(~((a == b) | (np.isnan(a) & np.isnan(b)))))
