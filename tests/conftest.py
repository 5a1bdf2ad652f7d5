import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def fe():
    """One libssfe context for the GPU tests (fails loudly if the library or the GPU is missing)."""
    from speechsplit_b200 import FrontEnd
    f = FrontEnd(0)
    yield f
    f.close()


@pytest.fixture(scope="session")
def fe_seq():
    """Context in the sequential filtfilt validation mode."""
    from speechsplit_b200 import FrontEnd, FrontEndConfig
    f = FrontEnd(0, FrontEndConfig(filtfilt_mode=1))
    yield f
    f.close()
