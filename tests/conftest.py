import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _cuda_available() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    from oracle import msed_oracle
    msed_oracle.build()
    return msed_oracle


@pytest.fixture(scope="session")
def msed_lib():
    """libmsed_b200.so, built if stale.  Never skipped on a GPU box: a missing library is a failure."""
    from mossco_code_b200 import _abi, build
    build.build()
    return _abi.load()


@pytest.fixture(scope="session")
def gpu(msed_lib):
    if not _cuda_available():
        pytest.fail("test marked gpu but no CUDA device is visible (the product has no CPU path)")
    return msed_lib
