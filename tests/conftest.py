import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Build the product library and the oracle once per session (nvcc/g++ work without a GPU)."""
    from oracle import pm_oracle
    from protocol_b200 import build as pm_build

    pm_build.build()
    pm_oracle.build()
    yield


def has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False
