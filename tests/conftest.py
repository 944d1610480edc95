import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def frontend():
    """One bce_gpu context for the whole GPU session (fails loudly without the CUDA library)."""
    from bce_b200 import Frontend
    fe = Frontend(0)
    yield fe
    fe.close()
