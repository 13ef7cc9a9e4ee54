import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# The XTC codec (libxdrfile) is third-party; the reference vendors it as a binary.  Where that tree is present (the build
# container, not the GPU box) the CPU tests of the traj/xtc mirror use it; elsewhere they skip.
_XDR = "/root/reference/traj/xtc/libxdrfile.so"
if "GOCHEM_XDRFILE" not in os.environ and os.path.exists(_XDR):
    os.environ["GOCHEM_XDRFILE"] = _XDR


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
