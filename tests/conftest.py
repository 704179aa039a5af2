import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    """The C restatement is test infrastructure: build it once per session if it is not there."""
    import oracle
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        oracle.build()
    yield
