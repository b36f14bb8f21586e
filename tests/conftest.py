import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_DATA = os.path.join(GOLDEN, "ref")             # copies of the reference's swga/tests/data fixtures


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import swga_oracle
    swga_oracle.lib()
    return swga_oracle


@pytest.fixture(scope="session")
def dsk_golden():
    import json
    with open(os.path.join(GOLDEN, "dsk_golden.json")) as f:
        return json.load(f)
