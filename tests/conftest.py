import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _fresh_library_state(request):
    """Raw-kernel tests must not inherit process-global library state from an earlier test: the dropout step
    counter a trainer registered (ergm_set_rng_step_ptr) is cleared before every GPU test."""
    if request.node.get_closest_marker("gpu") is not None:
        import torch
        if torch.cuda.is_available():
            from ergm_b200 import _lib
            _lib.lib().ergm_set_rng_step_ptr(None)
    yield
