import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def fake_ops(monkeypatch):
    """The emulated backend (tests/fake_backend.py) in place of the C ABI: host logic without a GPU."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import fake_backend
    fake_backend.install(monkeypatch)
    return fake_backend
