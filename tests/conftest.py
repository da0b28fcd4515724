import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ifl():
    """The product package (host mirror + C-ABI loader)."""
    import incfluid_b200
    incfluid_b200.build()
    return incfluid_b200


@pytest.fixture(scope="session")
def abi(ifl):
    return ifl.abi


@pytest.fixture(scope="session")
def oracle_cls():
    from oracle_harness import OracleSolver, build_oracle
    build_oracle()
    return OracleSolver
