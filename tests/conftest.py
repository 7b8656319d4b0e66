import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference (oracle/_ref/libgfsref.so); skipped where it cannot be had."""
    from oracle.pyoracle import Reference
    try:
        return Reference()
    except (FileNotFoundError, OSError) as e:
        pytest.skip("reference build unavailable: %s" % e)
