import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def refcv():
    """The unmodified reference through oracle/_ref/libhsref_cv.so (skips when it was not built)."""
    from oracle import pyoracle
    if not pyoracle.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return pyoracle.RefCV


@pytest.fixture(scope="session")
def gpu_ctx():
    from hairsplitter_b200 import api
    ctx = api.Context(0)  # raises when there is no GPU: the -m gpu tests must fail loudly, not skip
    yield ctx
    ctx.close()
