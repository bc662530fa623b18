"""pytest configuration: registers the ``gpu`` marker and puts the repo root on ``sys.path``.

``-m "not gpu"``: oracle vs golden vectors, host logic, C-ABI symbols (no GPU needed).
``-m gpu``: parity of the CUDA path (through the C ABI) against the oracle, on a B200.
"""
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


@pytest.fixture(scope="session")
def golden():
    def _load(name):
        return np.load(os.path.join(GOLDEN, name))
    return _load


def has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine on cuda:0; a gpu-marked test that cannot get one FAILS (no silent fallback)."""
    import python_motionplanning_b200 as mp
    return mp.Engine(0)


# parity contract of BASELINE.json north_star: |gpu - ref| <= 1e-9 * max(|ref|, 1) on every state component
REL_TOL_F64 = 1e-9


def rel_err(a, ref, floor=1.0):
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return np.abs(a - ref) / np.maximum(np.abs(ref), floor)
