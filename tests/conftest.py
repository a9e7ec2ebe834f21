import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


@pytest.fixture(scope="session")
def citeseer():
    from tests.fixtures import load_fixture
    return load_fixture("citeseer")


@pytest.fixture(scope="session")
def mico():
    from tests.fixtures import load_fixture
    return load_fixture("mico")
