import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def rowmax_rel(a, b):
    """max |a-b| / max|b| along the last axis - the comparison of SURVEY.md 7.3: rtol plus an atol
    of tol*max|row| so that exact zero crossings do not produce spurious failures."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = np.abs(b).max(axis=-1, keepdims=True)
    scale = np.where(scale == 0, 1.0, scale)
    return float(np.max(np.abs(a - b) / scale))


@pytest.fixture(scope="session")
def golden2():
    return dict(np.load(os.path.join(GOLDEN, "config2_chain.npz")))


@pytest.fixture(scope="session")
def golden_nl2():
    return dict(np.load(os.path.join(GOLDEN, "nl2_resum.npz")))


@pytest.fixture(scope="session")
def fftlog_kat():
    return dict(np.load(os.path.join(GOLDEN, "fftlog_kat.npz")))


@pytest.fixture(scope="session")
def fiber_kat():
    return dict(np.load(os.path.join(GOLDEN, "fiber_kat.npz")))


@pytest.fixture(scope="session")
def golden_opts():
    """reference outputs for optiresum / IRcutoff / LambdaIR variants (tests/golden/make_golden_options.py)"""
    return dict(np.load(os.path.join(GOLDEN, "options_resum.npz")))


@pytest.fixture(scope="session")
def interp_kat():
    """reference outputs for the un-binned likelihood products (tests/golden/make_golden_interp.py)"""
    return dict(np.load(os.path.join(GOLDEN, "interp_kat.npz")))
