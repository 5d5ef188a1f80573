import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref (the compiled reference)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    """The checker libraries are built on demand (liboracle.so always; _ref only where the
    reference tree is mounted)."""
    from oracle import oracle as O

    if not os.path.exists(O.LIB_ORACLE):
        O.build(ref=True)
    yield
