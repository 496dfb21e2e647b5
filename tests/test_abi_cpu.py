"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/bgg.h declares, and
refuses to run without a CUDA device (no CPU fallback).  No compute calls are made here."""
import ctypes
import os
import re

import pytest

import common  # noqa: F401  (sets sys.path)
import bgg_b200 as bg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "bgg.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bgg_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(bg.LIB_PATH):
        pytest.fail(f"{bg.LIB_PATH} has not been built (python -c 'import __graft_entry__ as g; g.build()')")
    lib = ctypes.CDLL(bg.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/bgg.h but not exported"
    assert sorted(bg.exported_symbols()) == declared, "bgg_b200.exported_symbols() is out of date with include/bgg.h"


def test_instance_pod_layout_matches_numpy_view():
    lib = ctypes.CDLL(bg.LIB_PATH)
    lib.bgg_instance_bytes.restype = ctypes.c_size_t
    assert lib.bgg_instance_bytes() == bg.INSTANCE_DTYPE.itemsize


def test_no_cpu_fallback():
    lib = bg.lib()
    lib.bgg_device_count.restype = ctypes.c_int
    if lib.bgg_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(bg.BggError, match="no CUDA device"):
        bg.BatchedMPC(20, 0.05, common.wl.robot())
