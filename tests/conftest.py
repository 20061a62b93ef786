import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the library and the oracle once (nvcc cross-compiles without a GPU)."""
    from mllp_b200 import build as b
    b.build()
    from oracle import pdhg_oracle
    pdhg_oracle.build()


SMALL = ["afiro", "sc50a", "sc105", "adlittle", "blend", "share2b", "kb2"]
MID = ["25fv47", "pilot87", "d2q06c", "dfl001"]
LARGE = ["ken-18", "osa-60", "pds-20"]
