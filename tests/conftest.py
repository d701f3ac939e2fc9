import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def cref():
    """The C oracle (test infrastructure), built on demand."""
    from oracle import cref as c
    c.lib()
    return c


@pytest.fixture(scope='session')
def dg():
    """The product library on cuda:0.  No fallback: a missing library or GPU is a hard failure."""
    from crypto_b200 import lib
    lib.init(0)
    return lib
