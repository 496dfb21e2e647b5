"""Binding over controller::MPCController of the C++ host layer (host/mpc_controller_b200.{h,cpp}, libmpc_b200.so): the MPC
thread of the reference's controller (controllers/mpc_controller.cpp:286-399, MPCUpdate, and :518-573, GaitOpt) for a whole
batch, one call per pass of its loop body -- bgg_controller_tick_batch underneath: a single launch sequence on the stream and a
single read-back per tick.  The schedule (which of the three modes a pass takes, deriv_ready per robot, the cost reduction) is
decided in the C++ class; this file only moves arrays."""
import ctypes as C
import os

import numpy as np

import bgg_b200 as bg

LS_SIZE = 10   # gait_optimizer.h: LS_SIZE
MODES = {0: "solve", 1: "solve_and_gait_opt", 2: "line_search"}
_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
_lib = None


def _shim():
    global _lib
    if _lib is None:
        bg.lib()   # libbgg_b200.so first: the shim links against it
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmpc_b200.so")
        if not os.path.exists(path):
            raise bg.BggError(f"{path} is missing: run bilevel-gait-gen_b200/build.sh")
        L = C.CDLL(path)
        L.bggc_create.restype = C.c_void_p
        L.bggc_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.bggc_destroy.argtypes = [C.c_void_p]
        L.bggc_next_mode.argtypes = [C.c_void_p]
        L.bggc_run_num.argtypes = [C.c_void_p]
        L.bggc_mpc_update.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.bggc_set_initial_config.argtypes = [C.c_void_p, _dp]
        L.bggc_targets_from_traj.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _ip]
        L.bggc_results.argtypes = [C.c_void_p, _ip, _ip, _dp, _dp, _dp, _ip, _dp, _ip, _dp, _ip]
        _lib = L
    return _lib


class MPCController:
    def __init__(self, mpc, gait_opt_freq, ls_size=LS_SIZE):
        self.mpc, self.ls_size = mpc, int(ls_size)
        self.L = _shim()
        self.c = self.L.bggc_create(mpc.h, mpc.B, mpc.N, int(gait_opt_freq), self.ls_size)
        if not self.c:
            raise bg.BggError("MPCController: bad arguments")
        self.lp = None

    def __del__(self):
        if getattr(self, "c", None):
            self.L.bggc_destroy(self.c)
            self.c = None

    def SetInitialConfig(self, q):
        """q_des_ before the first targets call (mpc_controller.cpp:50-52): [B][19]."""
        q = np.ascontiguousarray(q, np.float64).reshape(self.mpc.B, 19)
        self.L.bggc_set_initial_config(self.c, q.ctypes.data_as(_dp))

    def GetTargetsFromTraj(self, time):
        """MPCController::GetTargetsFromTraj (mpc_controller.cpp:414-511) for the whole batch; needs mpc.SetKinematics()."""
        B = self.mpc.B
        t = np.ascontiguousarray(np.broadcast_to(np.asarray(time, np.float64), (B,)))
        q, v, f, st = np.zeros((B, 19)), np.zeros((B, 18)), np.zeros((B, 4, 3)), np.zeros(B, np.int32)
        if self.L.bggc_targets_from_traj(self.c, t.ctypes.data_as(_dp), q.ctypes.data_as(_dp), v.ctypes.data_as(_dp), f.ctypes.data_as(_dp),
                                         st.ctypes.data_as(_ip)) != 0:
            raise bg.BggError(bg.lib().bgg_last_error().decode())
        return dict(q_des=q, v_des=v, force_des=f, status=st)

    @property
    def run_num(self):
        return self.L.bggc_run_num(self.c)

    def mode(self):
        """'line_search' | 'solve_and_gait_opt' | 'solve' for the coming tick (mpc_controller.cpp:323-345)."""
        return MODES[self.L.bggc_next_mode(self.c)]

    def MPCUpdate(self, state, time, ee_locations):
        """One pass of the while-loop body for the whole batch.  Returns dict(mode, status, cost, ...)."""
        B, K = self.mpc.B, self.ls_size
        s = np.ascontiguousarray(state, np.float64)
        t = np.ascontiguousarray(np.broadcast_to(np.asarray(time, np.float64), (B,)))
        e = np.ascontiguousarray(ee_locations, np.float64)
        m = self.L.bggc_mpc_update(self.c, s.ctypes.data_as(_dp), t.ctypes.data_as(_dp), e.ctypes.data_as(_dp))
        if m < 0:
            raise bg.BggError(bg.lib().bgg_last_error().decode())
        st, it, rd, best = np.zeros(B, np.int32), np.zeros(B, np.int32), np.zeros(B, np.int32), np.zeros(B, np.int32)
        al, co, cr = np.zeros(B), np.zeros(B), np.zeros(B)
        dH, lc, lq = np.zeros((B, 4, bg.MAX_CONTACTS)), np.zeros((B, K)), np.zeros((B, K), np.int32)
        self.L.bggc_results(self.c, st.ctypes.data_as(_ip), it.ctypes.data_as(_ip), al.ctypes.data_as(_dp), co.ctypes.data_as(_dp),
                            cr.ctypes.data_as(_dp), rd.ctypes.data_as(_ip), dH.ctypes.data_as(_dp), best.ctypes.data_as(_ip),
                            lc.ctypes.data_as(_dp), lq.ctypes.data_as(_ip))
        self.deriv_ready = rd.astype(bool)
        self.cost_red = cr
        res = dict(mode=MODES[m], status=st, iters=it, alpha=al, cost=co, best=best)
        if MODES[m] == "line_search":
            res.update(ls_costs=lc, quality=lq)
        if MODES[m] == "solve_and_gait_opt":
            self.lp = self.mpc.controller_step()
            counts = self.mpc.GetContactTimes()[2]
            res.update(grad_status=np.where(rd != 0, 0, 1), dHdtheta=[np.concatenate([dH[b, e, :counts[b, e]] for e in range(4)]) for b in range(B)])
        return res
