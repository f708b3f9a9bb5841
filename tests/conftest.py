import os, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def pytest_collection_modifyitems(config, items):
    """Plain `pytest tests` on a machine without CUDA: skip the gpu-marked tests instead of failing in the driver.
    (With CUDA present nothing is skipped: a missing libsvae_b200.so must fail loudly there.)"""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run with -m gpu under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
