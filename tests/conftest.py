import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs the compiled reference under oracle/_ref (built where /root/reference exists)")


def _have_ref():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_replay.so"))


@pytest.fixture(scope="session")
def port():
    from oracle.portlib import PortLib
    return PortLib()


@pytest.fixture(scope="session")
def ref():
    if not _have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    from oracle.reflib import RefLib
    return RefLib("replay")


@pytest.fixture(scope="session")
def ref_native():
    if not _have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    from oracle.reflib import RefLib
    return RefLib("native")
