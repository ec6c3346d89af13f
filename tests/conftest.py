import os
import sys

import pytest

# The emulated-rank tests (tests/test_gpu_dp.py) run kernels of several "ranks" of ONE process that wait for each other;
# lazy module loading may synchronise the context at a kernel's first launch, so load everything up front.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def ctx():
    """Device context; GPU tests fail loudly (not skip) when the library or the GPU is missing."""
    from simplesr_b200 import model_builder
    return model_builder.get_context(0)
