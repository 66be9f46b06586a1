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
    from oracle import fqz_oracle

    fqz_oracle.lib()
    return fqz_oracle


@pytest.fixture(scope="session")
def sample_fq():
    with open(os.path.join(ROOT, "tests", "golden", "sample.fq"), "rb") as f:
        return f.read()
