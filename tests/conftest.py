import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "gpu-montecarlo-risk_b200"))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def engine():
    """One engine on cuda:0 for the whole GPU session.  Fails (not skips) when the CUDA library or
    device is missing: GPU tests must never pass on a fallback."""
    from mcr_b200 import api
    e = api.Engine(0)
    yield e
    e.close()
