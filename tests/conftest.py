import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(HERE, "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def cuda_lib():
    """liboblas_b200.so initialised on cuda:0.  No fallback: a missing library or a
    failing init is an error, not a skip."""
    from oblas_b200 import capi
    lib = capi.lib()
    capi.init(0)
    return lib
