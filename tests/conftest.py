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


def pytest_sessionstart(session):
    """A fresh checkout has no built artefacts (they are git-ignored): build the library and the oracle once, exactly as
    `__graft_entry__.build()` does.  Nothing is built when both shared objects are already there."""
    lib = os.path.join(ROOT, "kornia-rs_b200", "lib", "libkornia_b200.so")
    ora = os.path.join(ROOT, "oracle", "libkornia_oracle.so")
    if not (os.path.exists(lib) and os.path.exists(ora)):
        import __graft_entry__

        __graft_entry__.build()


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
