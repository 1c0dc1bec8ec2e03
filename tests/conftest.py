import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # Some test modules import the ctypes binding at collection time: in a fresh checkout (built artefacts are not tracked)
    # build before collecting.  A no-op when everything is up to date; on the GPU box the prebuilt files travelled.
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def built():
    """Make sure the native pieces exist (no-op when they are already up to date)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def orc(built):
    from oracle import loader
    return loader.port()


@pytest.fixture(scope="session")
def refo(built):
    from oracle import loader
    r = loader.ref()
    if r is None:
        pytest.skip("oracle/_ref not built (no /root/reference in this environment)")
    return r


@pytest.fixture(scope="session")
def checker(built):
    """The strongest CPU checker available: the compiled reference if present, else the port."""
    from oracle import loader
    return loader.ref() or loader.port()


@pytest.fixture(scope="session")
def gpu(built):
    import libav_b200._lib as L
    L.init(0)
    return L
