"""Batched restatement of controller::MPCController's MPC thread (controllers/mpc_controller.cpp:286-399, MPCUpdate, and
:518-573, GaitOpt) over bgg_b200.BatchedMPC: one tick = one call, every instance of the batch advanced by the same
three-mode schedule the reference runs for its single robot,

    run_num % gait_opt_freq == 0 and the derivative is ready   ->  GaitOptimizer::LineSearch          (:323-336)
    (run_num + 1) % gait_opt_freq == 0                           ->  GetRealTimeUpdate, then GaitOpt    (:337-340)
    otherwise                                                    ->  GetRealTimeUpdate                  (:341-345)

with `deriv_ready` kept per instance (GaitOpt returns false when the last solve was not `Solved`, mpc.cpp:1047-1057).
In a line-search tick the instances whose derivative is not ready get a zero step: all of their candidates are the
unchanged contact schedule, i.e. the plain real-time update the reference would run for them.

Host-side pieces of the reference loop that are not on the device path stay with the caller: the mutex-protected state
hand-off (:304-317), AdjustForCurrentContacts (the C++ shim's MPC::AdjustForCurrentContacts), visualisation and the
statistics log (host/mpc_b200.cpp: PrintStatLineToFile)."""
import numpy as np

LS_SIZE = 10   # gait_optimizer.h: LS_SIZE


class MPCController:
    def __init__(self, mpc, gait_opt_freq, ls_size=LS_SIZE):
        self.mpc = mpc
        self.gait_opt_freq = int(gait_opt_freq)
        self.ls_size = int(ls_size)
        self.run_num = 0
        B = mpc.B
        self.deriv_ready = np.zeros(B, bool)
        self.prev_cost = np.full(B, 1e10)     # mpc_controller.cpp:293
        self.cost_red = np.zeros(B)
        self.lp = None

    def mode(self):
        """'line_search' | 'solve_and_gait_opt' | 'solve' for the coming tick (mpc_controller.cpp:323-345)."""
        r, f = self.run_num, self.gait_opt_freq
        if r % f == 0 and r > 0 and self.deriv_ready.any():
            return "line_search"
        if (r + 1) % f == 0 and r > 0:
            return "solve_and_gait_opt"
        return "solve"

    def MPCUpdate(self, state, time, ee_locations):
        """One pass of the while-loop body for the whole batch.  Returns dict(mode, status, cost, best)."""
        mpc, B = self.mpc, self.mpc.B
        mode = self.mode()
        res = dict(mode=mode, best=np.full(B, -1, np.int32))
        if mode == "line_search":
            step = self.lp["step"].copy()
            step[~self.deriv_ready] = 0.0
            ls = mpc.LineSearch(state, time, ee_locations, self.lp["xk"], step, K=self.ls_size)
            res.update(best=ls["best"], ls_costs=ls["costs"], quality=ls["quality"])
            self.deriv_ready[:] = False
        else:
            out = mpc.GetRealTimeUpdate(state, time, ee_locations)
            res.update(status=out["status"], cost=out["cost"], alpha=out["alpha"], iters=out["iters"])
            self.cost_red = self.prev_cost - out["cost"]
            self.prev_cost = out["cost"].copy()
            if mode == "solve_and_gait_opt":   # MPCController::GaitOpt, :518-573
                g = mpc.ComputeCostFcnDerivWrtContactTimes()
                self.lp = mpc.OptimizeContactTimes(time)
                self.deriv_ready = g["status"] == 0
                res.update(grad_status=g["status"], dHdtheta=g["dHdtheta"])
            else:
                self.deriv_ready[:] = False
        self.run_num += 1
        return res
