import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long CPU check, opt in with EAQHM_SLOW=1")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("EAQHM_SLOW") == "1":
        return
    skip = pytest.mark.skip(reason="set EAQHM_SLOW=1 to run")
    for it in items:
        if "slow" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def lib():
    from eaqhm_b200 import _lib
    _lib.build()
    return _lib.lib()
