import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def hh():
    import hedgehog_jl_b200 as m
    return m


@pytest.fixture(scope="session")
def oracle():
    """CPU checker (oracle/); never on the product path."""
    from oracle import oracle as O
    return O.OracleEngine()


@pytest.fixture(scope="session")
def cuda(hh):
    """The product engine. GPU tests must FAIL (not skip) if the CUDA library cannot run."""
    return hh.default_engine(0)
