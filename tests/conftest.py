import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def engine():
    """One hmsg_ctx on cuda:0 for the whole GPU session; fails loudly without library/GPU."""
    from holoagent_b200 import build as _b
    from holoagent_b200.engine import HmsgEngine
    if _b.needs_build() and os.path.exists("/usr/local/cuda/bin/nvcc"):
        _b.build()
    return HmsgEngine(0)
