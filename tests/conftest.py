import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port():
    from oracle import oracle
    oracle.build()
    return oracle.Port()


@pytest.fixture(scope="session")
def ref():
    """The compiled reference; prebuilt oracle/_ref travels to the GPU box."""
    from oracle import oracle
    try:
        return oracle.Ref()
    except (FileNotFoundError, OSError):
        oracle.build()
        try:
            return oracle.Ref()
        except (FileNotFoundError, OSError):
            pytest.skip("oracle/_ref not available (no /root/reference and no prebuilt .so)")


@pytest.fixture(scope="session")
def checker(port):
    """Best available oracle: compiled reference, else the port."""
    from oracle import oracle
    return oracle.load()
