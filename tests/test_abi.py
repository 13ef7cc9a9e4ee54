"""CPU tests of the boundary: the shared library loads and exports every symbol include/gochem_b200.h
declares; without a GPU compute calls fail loudly instead of falling back."""
import ctypes as C
import os

import numpy as np
import pytest

from gochem_b200 import _lib


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _lib.declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gochem_b200.h but not exported"
    assert lib.gcb_version() >= 100


def test_no_oracle_in_product():
    """The product tree must not reference oracle/ (no CPU fallback)."""
    root = os.path.dirname(os.path.abspath(_lib.__file__))
    for dp, _dn, fn in os.walk(root):
        if "build" in dp.split(os.sep) or "__pycache__" in dp:
            continue
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liborc" not in txt, f


@pytest.mark.skipif(_lib.device_count() > 0, reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.gcb_traj_create(0, 10, 4, C.byref(h))
    assert rc == _lib.ERR_CUDA
    assert len(_lib.last_error()) > 0
    with pytest.raises(_lib.GcbError):
        _lib.DeviceTraj(10, 4)


def test_argument_errors_do_not_need_a_gpu():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.gcb_traj_create(0, 0, 4, C.byref(h)) == _lib.ERR_SHAPE
    assert lib.gcb_set_kernel_variant(99) == _lib.ERR_ARG
    assert lib.gcb_set_kernel_variant(_lib.KERNEL_AUTO) == 0
    assert lib.gcb_traj_natoms(None) == -1
    assert lib.gcb_sync(None) == _lib.ERR_ARG
