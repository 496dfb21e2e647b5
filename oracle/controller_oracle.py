"""TEST INFRASTRUCTURE -- CPU restatement of controller::MPCController's MPC loop for ONE robot, on top of the oracle MPC
(pyoracle.OracleMPC) and the gait-optimiser restatement (gait_oracle.py).  Follows controllers/mpc_controller.cpp:286-399
(MPCUpdate: the three-mode schedule) and :518-573 (GaitOpt), gait_optimizer.cpp:671-753 (LineSearch: arg-min over
LS_SIZE copies, index 0 when every copy is primal infeasible, then SetWarmStartTrajectory(best copy)).
Only tests/ may import this file; the product path is bilevel-gait-gen_b200/mpc_controller.py over the CUDA library."""
import numpy as np

import gait_oracle as go


class ControllerOracle:
    def __init__(self, o, gait_opt_freq, ls_size=10):
        self.o, self.freq, self.ls_size = o, int(gait_opt_freq), int(ls_size)
        self.run_num, self.deriv_ready = 0, False
        self.xk = self.step = None

    def mode(self):
        r, f = self.run_num, self.freq
        if r % f == 0 and r > 0 and self.deriv_ready:
            return "line_search"
        if (r + 1) % f == 0 and r > 0:
            return "solve_and_gait_opt"
        return "solve"

    def tick(self, state, time, ee, step_override=None):
        o, mode = self.o, self.mode()
        res = dict(mode=mode, best=-1)
        if mode == "line_search":
            ct = go.contact_times(o)
            step = self.step if step_override is None else step_override
            best, costs, q = go.line_search(o, state, time, ee, ct, self.xk, step, ls_size=self.ls_size)
            res.update(best=best, ls_costs=costs, quality=q)
            idx = best if best >= 0 else 0
            times = go.contact_times_for(ct, self.xk, step, idx / self.ls_size)
            for e in range(4):
                o.set_contact_times(e, times[e])
            o.solve(state, time, ee, real_time=True)   # the parent adopts the winning copy's trajectory
            self.deriv_ready = False
        else:
            st = o.solve(state, time, ee, real_time=True)
            res.update(status=st, cost=o.stats()["cost"])
            if mode == "solve_and_gait_opt":
                terms = go.derivative_terms(o) if st == 0 else None
                if terms is None:
                    self.deriv_ready = False
                else:
                    ct = go.contact_times(o)
                    g = go.cost_gradient(o, terms)
                    self.step = go.solve_gait_lp(ct, g, time)
                    self.xk = np.concatenate([t for t, _ in ct])
                    self.deriv_ready = True
                    res.update(dHdtheta=g)
            else:
                self.deriv_ready = False
        self.run_num += 1
        return res
