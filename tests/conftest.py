import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def b200():
    import _pkg
    return _pkg.load()


@pytest.fixture(scope="session")
def oracle():
    from oracle import aekl_ref
    return aekl_ref
