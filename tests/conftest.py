import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")
    # a fresh checkout has no built artefacts (they are git-ignored): build them once, as __graft_entry__.build() does
    if not (os.path.exists(os.path.join(ROOT, "graphminer_b200", "libgminer_b200.so"))
            and os.path.exists(os.path.join(ROOT, "oracle", "libgm_oracle.so"))):
        import subprocess
        subprocess.check_call(["make", "-s", "-j8", "-C", ROOT, "all"])


@pytest.fixture(scope="session")
def citeseer():
    from tests.fixtures import load_fixture
    return load_fixture("citeseer")


@pytest.fixture(scope="session")
def mico():
    from tests.fixtures import load_fixture
    return load_fixture("mico")
