"""Host logic of the batched MPCController schedule (bilevel-gait-gen_b200/mpc_controller.py) against the reference's
mode rules (controllers/mpc_controller.cpp:323-345), with a stand-in for the CUDA-backed BatchedMPC -- no GPU needed."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "bilevel-gait-gen_b200"))
import mpc_controller as mc   # noqa: E402


class FakeMPC:
    def __init__(self, B, grad_ok):
        self.B, self.grad_ok, self.calls = B, np.asarray(grad_ok), []

    def GetRealTimeUpdate(self, s, t, e):
        self.calls.append("solve")
        return dict(status=np.zeros(self.B, np.int32), cost=np.full(self.B, 10.0 - len(self.calls)), alpha=np.ones(self.B), iters=np.ones(self.B, np.int32))

    def ComputeCostFcnDerivWrtContactTimes(self):
        self.calls.append("grad")
        return dict(status=np.where(self.grad_ok, 0, 1).astype(np.int32), dHdtheta=[np.zeros(3)] * self.B)

    def OptimizeContactTimes(self, t):
        self.calls.append("lp")
        return dict(step=np.ones((self.B, 4, 12)), xk=np.zeros((self.B, 4, 12)))

    def LineSearch(self, s, t, e, xk, step, K=10):
        self.calls.append("ls")
        self.last_step = step.copy()
        return dict(best=np.zeros(self.B, np.int32), costs=np.zeros((self.B, K)), quality=np.zeros((self.B, K), np.int32))


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


def test_mode_sequence_matches_the_reference_chain():
    for freq in (2, 3, 5):
        fake = FakeMPC(3, [True, True, True])
        c = mc.MPCController(fake, gait_opt_freq=freq)
        got = [c.MPCUpdate(None, 0.0, None)["mode"] for _ in range(13)]
        assert got == reference_modes(freq, True, 13)


def test_instances_without_a_derivative_get_a_zero_step():
    fake = FakeMPC(3, [True, False, True])
    c = mc.MPCController(fake, gait_opt_freq=2)
    modes = [c.MPCUpdate(None, 0.0, None)["mode"] for _ in range(3)]
    assert modes == ["solve", "solve_and_gait_opt", "line_search"]
    assert np.all(fake.last_step[1] == 0.0) and np.all(fake.last_step[0] == 1.0) and np.all(fake.last_step[2] == 1.0)
    assert not c.deriv_ready.any()
    # nobody ready -> the tick is a plain solve, as in the reference
    fake2 = FakeMPC(2, [False, False])
    c2 = mc.MPCController(fake2, gait_opt_freq=2)
    assert [c2.MPCUpdate(None, 0.0, None)["mode"] for _ in range(3)] == reference_modes(2, False, 3)


def test_cost_reduction_bookkeeping():
    fake = FakeMPC(2, [True, True])
    c = mc.MPCController(fake, gait_opt_freq=4)
    c.MPCUpdate(None, 0.0, None)
    first = c.prev_cost.copy()
    c.MPCUpdate(None, 0.0, None)
    assert np.allclose(c.cost_red, first - c.prev_cost)   # cost_red = prev_cost - GetCost(), mpc_controller.cpp:372
