import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree C-ABI library (built by __graft_entry__.build())."""
    from cfen_vit_dehazing_b200.csrc import build as _b  # noqa: F401
    return _b.build()
