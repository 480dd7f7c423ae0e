import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REF = "/root/reference"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


needs_reference = pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference not present")


@pytest.fixture(scope="session")
def golden_names():
    return sorted(f[:-6] for f in os.listdir(GOLDEN) if f.endswith(".tlagz"))
