import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def gpu_engine():
    """The engine on a real device.  A machine WITH a GPU where the library
    does not load or sees no sm_100 device is a failure, not a skip."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device in this container")
    from contractn_b200 import engine

    lib = engine.load_library()
    assert lib.ctn_device_count() > 0, "libctn_b200.so loaded but found no compute-10.x device"
    return engine
