import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def abi_model(monkeypatch):
    """Swap the loaded CUDA library for the software model of its host entry points (tests/_abi_model.py) for ONE test,
    so that the Python host side above the C ABI can be exercised without a GPU.  Candidate generation on the device is
    switched off (the Optimizer then draws on the host, same bits).  Test infrastructure: the product never does this."""
    import skopt_b200  # noqa: F401
    from skopt_b200 import _lib
    from skopt_b200.optimizer import optimizer as optimizer_module
    from tests._abi_model import AbiModel
    if os.environ.get("SKB_TESTS_ON_DEVICE") == "1":
        # on a B200 box: `SKB_TESTS_ON_DEVICE=1 python -m pytest tests/test_host_optimizer.py` runs the same host-level
        # tests against the real libskb.so (nothing is swapped; `on_device` lets a test skip its call-log assertions)
        class _RealLibrary:
            on_device, calls = True, ()
        return _RealLibrary()
    model = AbiModel()
    monkeypatch.setattr(_lib, "_lib", model)
    monkeypatch.setattr(optimizer_module, "device_dim_descs", lambda space: None)
    return model
