import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_cuda():
    """True when a CUDA device is visible (asked of the driver through torch; no context is created)."""
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` are skipped (not failed) on a machine without a CUDA device, so a plain `pytest tests`
    goes green on a CPU-only box; on the GPU box nothing is skipped."""
    if _have_cuda():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (there is no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _build_once():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "scema_b200", "libscema_hist.so")) or \
       not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        g.build()


@pytest.fixture(scope="session", autouse=True)
def built():
    _build_once()


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference header (oracle/_ref). Skips where it has not been built."""
    from oracle.pyoracle import Reference, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return Reference()


@pytest.fixture(scope="session")
def hc():
    import scema_b200
    h = scema_b200.HistCluster(0)
    yield h
    h.close()
