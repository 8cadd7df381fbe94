import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.dirname(os.path.abspath(__file__))
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture
def oracle_backend(monkeypatch):
    """
    Swap the device GP for the CPU oracle so the HOST logic (array building, optimiser flow,
    sharding, blending) can be exercised without a GPU.  Test-only: the product has no such
    switch.
    """
    import oracle_george
    import emulator_b200.george_compat as gc

    monkeypatch.setattr(gc, "GP", oracle_george.OracleBackedGP)
    monkeypatch.setattr(gc, "pinned_evaluator", oracle_george.pinned_evaluator)
    return oracle_george.OracleBackedGP
