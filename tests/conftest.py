"""pytest configuration: marker registration and shared fixtures."""

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the live reference tree at /root/reference")


@pytest.fixture(scope="session")
def reference():
    """The live reference package, or skip when /root/reference is absent
    (e.g. on the GPU box)."""
    from oracle import ref_shim

    if not ref_shim.available():
        pytest.skip("reference tree not present")
    return ref_shim.load()
