"""pytest configuration: registers the `gpu` marker and common fixtures."""

import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_files(prefix):
    files = sorted(glob.glob(os.path.join(GOLDEN, f"{prefix}_*.npz")))
    assert files, f"no golden fixtures for {prefix}"
    return files


def load_cores(z, prefix):
    """Cores stored as <prefix>0..<prefix>{d-1} in a golden npz, reference-shaped."""
    out = []
    k = 0
    while f"{prefix}{k}" in z:
        out.append(np.array(z[f"{prefix}{k}"]))
        k += 1
    return out


@pytest.fixture(scope="session")
def oracle():
    from oracle import tt_oracle

    return tt_oracle
