"""Pins the oracle's restatement of the live RTI-MPC assembly (oracle/srb_mpc.cpp) on what the reference itself
fixes: problem dimensions (derived from mpc.cpp:1101-1127,1205-1214 and
mpc_single_rigid_body.cpp:323-341), the structural asserts inside the library (mpc.cpp:207,412;
mpc_single_rigid_body.cpp:296-297,441,473,885), the zero-dropping rule of utils/sparse_matrix_builder.cpp:25, and
self-consistency of the linearisation (C = f(x) - A x - B u reproduces the Euler step at the linearisation point).
No stored QP matrices exist in the reference (parity of values is GPU-vs-oracle, tests/test_gpu_parity.py)."""
import numpy as np
import pytest

import common
from common import wl
import pyoracle as po


@pytest.mark.parametrize("cfg_name,n,m_eq,m_ineq", [("a1_configuration", 372, 260, 752), ("a1_gait_opt_config", 732, 620, 1232)])
def test_problem_dimensions_at_t0(cfg_name, n, m_eq, m_ineq):
    cfg = wl.CONFIGS[cfg_name]
    o = common.make_oracle(cfg_name)
    o.assemble(np.asarray(cfg["srb_init"], float), 0.0, wl.EE_NOMINAL)
    sz = o.sizes()
    assert (sz["n"], sz["num_eq"], sz["num_ineq"], sz["m"]) == (n, m_eq, m_ineq, m_eq + m_ineq)
    assert (sz["nf"], sz["np"]) == (96, 24)
    assert sz["num_force_box"] == 160 and sz["num_cone"] == 320 and sz["num_start"] == 8 and sz["num_td"] == 0
    qp = o.qp()
    assert qp["A"].nnz <= 25000            # the reference's reserve, mpc.cpp:47
    assert qp["P"].nnz == n and np.all(qp["P"].diagonal() >= 1e-3)   # diagonal P, +1e-3 I (mpc.cpp:1090-1095)
    assert np.all(qp["A"].data != 0.0)     # exact zeros are never stored (sparse_matrix_builder.cpp:25)
    # first block row: -x_0 = -x_init
    A = qp["A"].toarray()
    assert np.array_equal(A[:12, :12], -np.eye(12)) and not A[:12, 12:].any()
    # the touch-down sample of every stance has all-zero spline weights -> empty force-box / cone rows
    fb = A[sz["num_dyn"]:sz["num_dyn"] + sz["num_force_box"]]
    assert (np.abs(fb).sum(1) == 0).sum() == 2 * 8
    # friction pyramid rows: 4 per sample, ub = 0
    r0 = sz["num_dyn"] + sz["num_force_box"]
    assert np.all(qp["ub"][r0:r0 + sz["num_cone"]] == 0)
    # foot-box rows touch exactly one state entry (-1 / +1) and 1-2 position variables
    r1 = r0 + sz["num_cone"]
    eb = A[r1:r1 + sz["num_ee_loc"]]
    nx = 12 * (cfg["num_nodes"] + 1)
    assert np.all((eb[:, :nx] != 0).sum(1) == 1)
    assert np.all(np.isin((eb[:, nx:] != 0).sum(1), (1, 2)))
    assert np.array_equal(eb[: len(eb) // 2], -eb[len(eb) // 2:])


def test_linearisation_reproduces_euler_step():
    cfg_name = "a1_configuration"
    cfg = wl.CONFIGS[cfg_name]
    states, _, ee = wl.batched_trot_inputs(cfg, 3, seed=2)
    for b in range(3):
        o = common.make_oracle(cfg_name, states[b])
        o.initial_run(states[b], ee[b])           # non-trivial forces and foot positions
        o.assemble(states[b], 0.0, ee[b])
        Ad, Bd, cd = o.node_dynamics()
        z = o.prev_qp_sol()
        N = cfg["num_nodes"]
        u = z[12 * (N + 1):]
        # x_k + dt f(x_k, t_k) == Ad x_k + Bd u + cd at the linearisation point: the defect of the previous trajectory
        # under the linearised dynamics equals its defect under the nonlinear Euler step (mpc.cpp:764-776)
        lin_defect = np.array([z[12 * (k + 1):12 * (k + 2)] - (Ad[k] @ z[12 * k:12 * (k + 1)] + Bd[k] @ u + cd[k]) for k in range(N)])
        merit_defect = (o.merit(z) - o.cost()) / 5000.0
        assert abs(np.abs(lin_defect).sum() - merit_defect) < 1e-9 * max(1.0, merit_defect)


def test_sizes_follow_the_sliding_horizon():
    """Knots are appended when the spline end falls inside the horizon and dropped once t0 passes a contact knot
    (trajectory.cpp:225-246); touch-down rows appear when the next touch-down is closer than 0.75 swing (mpc.cpp:1205-1214)."""
    cfg_name = "a1_configuration"
    cfg = wl.CONFIGS[cfg_name]
    o = common.make_oracle(cfg_name)
    init = np.asarray(cfg["srb_init"], float)
    o.initial_run(init, wl.EE_NOMINAL)
    seen_nu, seen_td = set(), set()
    for step in range(14):
        t0 = 0.05 * step
        ee_now = np.array([o.ee_at(e, t0) for e in range(4)])
        st = o.solve(init, t0, ee_now)
        assert st in (0, 1)
        sz = o.sizes()
        assert sz["n"] == 12 * 21 + sz["nf"] + sz["np"]
        assert sz["num_eq"] == 252 + 8 + sz["num_td"]
        seen_nu.add(sz["nf"] + sz["np"])
        seen_td.add(sz["num_td"])
        for e in range(4):
            f = o.foot(e)
            assert f.start_time() <= t0 + 1e-12 and f.end_time() >= t0 + 1.0 - 1e-9
    assert len(seen_nu) > 1 and seen_td >= {0, 4}


def test_quaternion_maps_round_trip():
    rng = np.random.default_rng(0)
    for scale in (1e-9, 1e-4, 1e-2, 0.3, 2.0):
        v = rng.normal(size=3)
        v *= scale / np.linalg.norm(v)
        q = po.quat_exp3(v)
        assert abs(np.linalg.norm(q) - 1) < 1e-8
        assert np.abs(po.quat_log3(q) - v).max() < 1e-9 * max(1.0, scale)


def test_infeasible_foot_start_widens_the_foot_box():
    """Feet pinned at the origin violate the foot box: the QP is primal infeasible, the previous solution is kept and
    the box grows by 5 cm per solve until the problem becomes feasible, then shrinks back
    (mpc_single_rigid_body.cpp:115-144, 929-937)."""
    cfg_name = "a1_configuration"
    cfg = wl.CONFIGS[cfg_name]
    o = common.make_oracle(cfg_name)
    init = np.asarray(cfg["srb_init"], float)
    bad_ee = np.zeros((4, 3))
    boxes, statuses = [], []
    for _ in range(10):   # |hip_x| = 0.2055 needs box/2 >= 0.2055: six widenings from 0.15
        statuses.append(o.solve(init, 0.0, bad_ee))
        boxes.append(o.stats()["ee_box_x"])
    assert statuses[0] == 3 and boxes[0] == pytest.approx(0.20)
    assert 0 in statuses
    assert max(boxes) > 0.15 and boxes[-1] <= max(boxes)
