import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    import pyoracle

    return pyoracle.load()


@pytest.fixture(scope="session")
def cuda():
    """the product API; fails loudly (no fallback) when the library or a device is missing"""
    import ipctk_b200

    return ipctk_b200.library()


@pytest.fixture(scope="session")
def scenes():
    import ipctk_b200

    return ipctk_b200._pkg.scenes
