import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle behind the 13-call boundary (test infrastructure)."""
    from strom_b200.beagle import BeagleAPI
    from strom_b200.build import build_oracle
    return BeagleAPI(build_oracle())


@pytest.fixture(scope="session")
def engine():
    """The product library.  Never falls back: a missing .so is an error."""
    from strom_b200 import beagle
    return beagle.engine()
