import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `pytest -m gpu`)")


_gpu = None


def have_gpu() -> bool:
    global _gpu
    if _gpu is None:
        try:
            from raytracinginaweekend_b200 import _cabi
            _gpu = _cabi.device_count() > 0
        except Exception:
            _gpu = False
    return _gpu


def pytest_collection_modifyitems(config, items):
    if have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device: the product path has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.lib()
    return binding


_worlds = {}


@pytest.fixture(scope="session")
def oracle_world(oracle):
    def get(name):
        if name not in _worlds:
            _worlds[name] = oracle.OracleWorld.named(name)
        return _worlds[name]
    return get


@pytest.fixture(scope="session", autouse=True)
def release_gpu_workspaces_at_the_end():
    yield
    if have_gpu():
        from raytracinginaweekend_b200 import _cabi
        _cabi.shutdown()
