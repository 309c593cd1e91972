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
    """BSIM-CMG 107 compiled for the host (oracle side).  Skips when neither the cached generated
    model nor the Verilog-A source is available."""
    from cedarsim.jl_b200 import models
    from cedarsim.jl_b200.va.build import build_host
    if not models.available():
        pytest.skip("BSIM-CMG source / cache not available")
    return build_host(models.bsimcmg107())
