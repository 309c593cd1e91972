import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def host_bsimcmg():
    """True when BSIM-CMG can be built for the host (oracle side): needs the cached generated
    models or the Verilog-A source.  Tests pass it as `host=` to the circuit builders."""
    from cedarsim.jl_b200 import models
    if not models.available():
        pytest.skip("BSIM-CMG source / cache not available")
    return True
