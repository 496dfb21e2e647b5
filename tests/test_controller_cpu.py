"""Host logic of controller::MPCController (bilevel-gait-gen_b200/host/mpc_controller_b200.cpp, through its C entry points in
libmpc_b200.so) against the reference's mode rules (controllers/mpc_controller.cpp:323-345) -- the schedule only, no device work
(AdvanceWithoutDevice), so no GPU is needed."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "bilevel-gait-gen_b200")
_ip = C.POINTER(C.c_int32)


def _shim():
    so = os.path.join(PKG, "libmpc_b200.so")
    if not os.path.exists(so):
        pytest.skip("libmpc_b200.so is not built (python -c 'import __graft_entry__ as g; g.build()')")
    C.CDLL(os.path.join(PKG, "libbgg_b200.so"), mode=C.RTLD_GLOBAL)
    L = C.CDLL(so)
    L.bggc_create.restype = C.c_void_p
    L.bggc_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.bggc_destroy.argtypes = [C.c_void_p]
    L.bggc_next_mode.argtypes = [C.c_void_p]
    L.bggc_run_num.argtypes = [C.c_void_p]
    L.bggc_advance_without_device.argtypes = [C.c_void_p, _ip]
    return L


MODES = {0: "solve", 1: "solve_and_gait_opt", 2: "line_search"}


def reference_modes(freq, deriv_ok, ticks):
    """The if / else-if / else chain of MPCUpdate for one robot."""
    out, ready = [], False
    for run_num in range(ticks):
        if run_num % freq == 0 and run_num > 0 and ready:
            out.append("line_search")
            ready = False
        elif (run_num + 1) % freq == 0 and run_num > 0:
            out.append("solve_and_gait_opt")
            ready = deriv_ok
        else:
            out.append("solve")
            ready = False
    return out


def _run(L, freq, deriv_ok, ticks):
    B = len(deriv_ok)
    c = L.bggc_create(C.c_void_p(1), B, 20, freq, 10)   # the handle is not touched without device work
    assert c
    ok = np.asarray(deriv_ok, np.int32)
    got = []
    for k in range(ticks):
        assert L.bggc_run_num(c) == k
        nxt = L.bggc_next_mode(c)
        assert L.bggc_advance_without_device(c, ok.ctypes.data_as(_ip)) == nxt
        got.append(MODES[nxt])
    L.bggc_destroy(c)
    return got


def test_mode_sequence_matches_the_reference_chain():
    L = _shim()
    for freq in (2, 3, 5):
        assert _run(L, freq, [1, 1, 1], 13) == reference_modes(freq, True, 13)


def test_line_search_runs_when_any_robot_has_a_derivative_and_not_otherwise():
    L = _shim()
    assert _run(L, 2, [1, 0, 1], 7) == reference_modes(2, True, 7)        # a batch searches as soon as one robot can
    assert _run(L, 2, [0, 0, 0], 7) == reference_modes(2, False, 7)       # no derivative anywhere: the line-search tick is a plain solve
    assert "line_search" not in _run(L, 3, [0, 0], 10)


def test_bad_arguments_are_refused():
    L = _shim()
    assert not L.bggc_create(None, 4, 20, 3, 10)
    assert not L.bggc_create(C.c_void_p(1), 0, 20, 3, 10)
    assert not L.bggc_create(C.c_void_p(1), 4, 20, 0, 10)
