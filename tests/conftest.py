import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The oracle is (re)built on demand; the CUDA library must already be in-tree (build() makes it)."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "_build/liboracle.so"], check=True)
    yield


@pytest.fixture(scope="session")
def nmpc():
    import nmpc_b200

    nmpc_b200.lib()
    return nmpc_b200


@pytest.fixture(scope="session")
def gpu(nmpc):
    if nmpc.device_count() <= 0:
        pytest.fail("a test marked gpu is running without a CUDA device")
    return nmpc
