import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure only); built on demand with the committed Makefile."""
    from oracle import pyoracle
    pyoracle.build()
    pyoracle.set_threads(min(os.cpu_count() or 1, 16))
    return pyoracle


@pytest.fixture(scope="session")
def gpu_ctx():
    """A zkb_ctx on cuda:0 — fails (does not skip) when the library or the device is missing."""
    from zk_stark_project_b200 import lib
    return lib.Context(0)
