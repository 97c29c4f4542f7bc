import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HAVE_REFERENCE = os.path.exists("/root/reference/src/ORBextractor.cpp")
HAVE_REF_LIB = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "liborbref_parity.so")) or HAVE_REFERENCE


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def port():
    import oracle
    oracle.build()
    return oracle.Port()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
