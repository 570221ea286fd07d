import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "code-adaptive-prob-ode-solvers_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The CUDA library is git-ignored (built in-tree): build it if a fresh checkout lacks it.
    nvcc cross-compiles sm_100a without a GPU; this is a build step, not a fallback."""
    lib = os.path.join(PKG, "libpn_b200.so")
    if not os.path.exists(lib):
        import importlib.util

        spec = importlib.util.spec_from_file_location("pn_b200_build", os.path.join(PKG, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()


@pytest.fixture(scope="session")
def goldens():
    """Reference artefacts extracted by tests/golden/make_golden.py (never reads /root/reference)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_goldens.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import pn_oracle

    pn_oracle.lib()
    return pn_oracle
