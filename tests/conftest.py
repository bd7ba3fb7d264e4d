import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


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
