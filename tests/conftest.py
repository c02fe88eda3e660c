import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from _oracle import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def refserial():
    from _oracle import RefSerial, have_ref

    if not have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return RefSerial()
