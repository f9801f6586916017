import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def native_lib():
    """The built C-ABI library (built on demand: nvcc cross-compiles without a GPU)."""
    from heybuddy_b200 import _native, build

    if not os.path.exists(_native.lib_path()):
        build.build(verbose=False)
    return _native.load()


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch.device("cuda:0")

# The speech-embedding artefact cannot be downloaded offline: tests run on the seeded random init on purpose.
os.environ.setdefault("HEYBUDDY_B200_ALLOW_RANDOM_INIT", "1")
