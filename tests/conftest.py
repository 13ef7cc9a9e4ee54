import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# The XTC codec (libxdrfile) is third-party; the reference vendors it as a binary.  Where that tree is present (the build
# container, not the GPU box) the tests of the traj/xtc mirror bind the real codec; elsewhere they bind tests/xdrshim/xdrshim.c,
# a stand-in with the same five entry points over uncompressed XTC records, so the reader -> pinned float32 AoS -> device chain
# is exercised on every host (no skip).
_XDR = "/root/reference/traj/xtc/libxdrfile.so"


def build_xdrshim() -> str:
    """Compile tests/xdrshim/xdrshim.c next to its source (gcc is in the image); returns the library path."""
    import subprocess

    src = os.path.join(ROOT, "tests", "xdrshim", "xdrshim.c")
    out = os.path.join(ROOT, "tests", "xdrshim", "libxdrshim.so")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        tmp = out + ".%d.tmp" % os.getpid()
        subprocess.check_call([os.environ.get("CC", "gcc"), "-O2", "-shared", "-fPIC", "-o", tmp, src])
        os.replace(tmp, out)
    return out


if "GOCHEM_XDRFILE" not in os.environ:
    os.environ["GOCHEM_XDRFILE"] = _XDR if os.path.exists(_XDR) else build_xdrshim()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
