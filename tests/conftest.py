import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def ob():
    """The CPU oracle binding (test infrastructure)."""
    import oracle_binding
    oracle_binding.build_oracle()
    return oracle_binding


@pytest.fixture(scope="session")
def pkg():
    """The product package (ada-ray-tracer_b200/), loaded through __graft_entry__.load_package()."""
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ge.PKG_DIR, "libgcore_b200.so")):
        ge.build()
    return ge.load_package()


@pytest.fixture(scope="session")
def oscene(ob):
    return ob.OracleScene()


@pytest.fixture()
def gscene(pkg):
    """Product scene initialised like Scene.Init of the reference's default build."""
    pkg.set_partition(0, 1)
    s = pkg.Scene()
    s.Init()
    return s
