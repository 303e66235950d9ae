import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


@pytest.fixture(scope="session", autouse=True)
def _native_library():
    """Build libgfb200.so if the sources are newer (nvcc cross-compiles without a GPU)."""
    from genesis_forge_b200 import build_native

    try:
        build_native.build()
    except Exception as e:  # pragma: no cover
        print(f"[conftest] native build skipped: {e}")
