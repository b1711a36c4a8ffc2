import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ob():
    import oracle_bindings

    oracle_bindings.lib()
    return oracle_bindings


@pytest.fixture(scope="session")
def ta():
    """The product package (ctypes over lib/libtrialign.so).  Fails loudly if the library is not built."""
    import trialign_b200

    trialign_b200.library()
    return trialign_b200


@pytest.fixture(scope="session")
def gpu_ta(ta):
    if ta.device_count() < 1:
        pytest.fail("no CUDA device visible: -m gpu tests must run on the GPU box")
    return ta
