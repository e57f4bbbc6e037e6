import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_available() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def scenes_dir():
    return os.path.join(ROOT, "scenes")


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def default_scene(orc, scenes_dir):
    return orc.load_rscn(os.path.join(scenes_dir, "default.rscn"))


@pytest.fixture(scope="session")
def benchmark_scene(orc, scenes_dir):
    return orc.load_rscn(os.path.join(scenes_dir, "benchmark.rscn"))


@pytest.fixture(scope="session")
def rb():
    import raydar_b200
    from raydar_b200 import build
    build.build()                       # no-op when libraydar_cuda.so is up to date
    raydar_b200.load_library()
    return raydar_b200


@pytest.fixture(scope="session")
def hs(rb):
    import hostsim_py
    hostsim_py.lib()
    return hostsim_py
