import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    import orc
    orc.lib()
    return orc


@pytest.fixture(scope="session")
def caller():
    import needlestack_b200 as ns
    c = ns.Caller(0)
    yield c
    c.close()
