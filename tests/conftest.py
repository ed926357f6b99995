import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "anon-aadhaar-halo2_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def zk():
    """The product library, bound to cuda:0.  Fails loudly when it cannot run."""
    import b200zk

    b200zk.init(0)
    return b200zk
