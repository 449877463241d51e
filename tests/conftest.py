import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.dirname(os.path.abspath(__file__)) not in sys.path:
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))   # tests/seal_format.py


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle4096():
    from oracle.bfv_oracle import Oracle
    return Oracle(4096, seed=4673838)


@pytest.fixture(scope="session")
def oracle8192():
    from oracle.bfv_oracle import Oracle
    return Oracle(8192, seed=4673838)
