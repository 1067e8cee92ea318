import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Build the CUDA library and the oracle once per session when they are missing (fresh clone); the built .so files
    are git-ignored but travel to the GPU box with the snapshot, so nothing is rebuilt there."""
    from ad_mpc_b200 import build as b
    if not os.path.exists(b.SO):
        b.build()
    from oracle import oracle as orc
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        orc.build()
    yield


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
