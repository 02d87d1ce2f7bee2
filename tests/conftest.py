"""pytest configuration: markers + import plumbing.

`-m "not gpu"` = oracle vs the reference's golden vectors, host logic, C-ABI symbol
checks (runs anywhere).  `-m gpu` = parity tests proper: the CUDA path called through
the C-ABI vs the oracle (needs a B200).
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o

    o.lib()
    return o


@pytest.fixture(scope="session")
def kb():
    """The product package (loads libkornia_b200.so; fails loudly if it is missing)."""
    import kornia_rs_b200

    return kornia_rs_b200
