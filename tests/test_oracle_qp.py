"""Pins the oracle's QP solvers on the reference's own cross-solver problem (test/mpc_test.cpp:857-904): the
3-variable QP that its Clarabel and OSQP interfaces must both solve to within 1e-4 of each other (:951-953) with
matching dx = Px + q (:955-958).  The exact optimum is computed here in closed form (two equalities leave one degree
of freedom, so the QP is a 1-D quadratic clipped to an interval)."""
import numpy as np
import scipy.sparse as sp

import common  # noqa: F401
import pyoracle as po

MARGIN = 1e-4   # test/mpc_test.cpp:917

P = np.diag([3.001, 4.0, 0.5])
q = np.array([0.1, 4.6, 2.0])
Aeq = np.array([[1.0, 1.0, 0.0], [1.3, 0.0, 0.2]])
beq = np.array([1.0, 3.0])
G = np.array([[-2.0, 0.0, 0.9], [1.0, 8.0, 5.0]])
g_lb = np.array([-2.0, -5.0])
g_ub = np.array([3.1, 13.3])


def exact_solution():
    xp = np.linalg.lstsq(Aeq, beq, rcond=None)[0]
    d = np.linalg.svd(Aeq)[2][-1]                      # null direction
    # minimise over t: 1/2 (xp+td)'P(xp+td) + q'(xp+td)
    t = -(d @ (P @ xp + q)) / (d @ P @ d)
    lo, hi = -np.inf, np.inf
    for row, l, u in zip(G, g_lb, g_ub):
        a, c = row @ d, row @ xp
        for bound, sign in ((u, 1), (l, -1)):          # sign*(a t + c) <= sign*bound
            if abs(a) < 1e-15:
                continue
            lim = (bound - c) / a
            if sign * a > 0:
                hi = min(hi, lim)
            else:
                lo = max(lo, lim)
    return xp + np.clip(t, lo, hi) * d


def test_three_variable_qp_admm_and_ipm_agree_with_exact():
    x_star = exact_solution()
    assert np.allclose(Aeq @ x_star, beq, atol=1e-12)
    # OSQP (two-sided) form, qp_data.cpp:200-289 with using_clarabel_ = false
    A2 = np.vstack([Aeq, G])
    l2 = np.concatenate([beq, g_lb])
    u2 = np.concatenate([beq, g_ub])
    st = po.default_admm_settings()   # osqp_interface.cpp:16-31
    # The reference runs OSQP with polish=true (:17), which refines the ADMM point to solver accuracy; polishing is not
    # restated, so the ADMM restatement is run to a tighter eps to meet the test's 1e-4 margin.
    st.eps_abs = st.eps_rel = 1e-6
    st.max_iter = 4000
    r_admm = po.admm_solve(sp.csc_matrix(P), q, sp.csc_matrix(A2), l2, u2, settings=st)
    assert r_admm["status"] == 0
    # Clarabel (one-sided) form: [eq ; G x <= ub ; -G x <= -lb], qp_data.cpp with using_clarabel_ = true
    A1 = np.vstack([Aeq, G, -G])
    b1 = np.concatenate([beq, g_ub, -g_lb])
    is_eq = np.array([1, 1, 0, 0, 0, 0], dtype=bool)
    r_ipm = po.ipm_solve(sp.csc_matrix(P), q, sp.csc_matrix(A1), b1, is_eq)
    assert r_ipm["status"] == 0
    # REQUIRE(osqp.GetSolveQuality() == clarabel.GetSolveQuality()) and primal agreement to MARGIN
    assert np.abs(r_admm["x"] - r_ipm["x"]).max() < MARGIN
    assert np.abs(r_admm["x"] - x_star).max() < MARGIN
    assert np.abs(r_ipm["x"] - x_star).max() < 1e-7
    # dx = P x + q agreement (Computedx, clarabel_interface.cpp:604-612)
    assert np.abs((P @ r_admm["x"] + q) - (P @ r_ipm["x"] + q)).max() < MARGIN
    # stationarity with the returned multipliers (sign conventions of both forms)
    assert np.abs(P @ r_ipm["x"] + q + A1.T @ r_ipm["y"]).max() < 1e-6
    assert np.abs(P @ r_admm["x"] + q + A2.T @ r_admm["y"]).max() < 1e-2
    assert np.all(r_ipm["y"][2:] >= -1e-12)
    # the one-sided and two-sided multipliers describe the same thing: y_two = lam_upper - lam_lower
    y_two = np.concatenate([r_ipm["y"][:2], r_ipm["y"][2:4] - r_ipm["y"][4:6]])
    assert np.abs(y_two - r_admm["y"]).max() < 5e-2 * max(1.0, np.abs(y_two).max())


def test_infeasible_qp_is_reported():
    A = sp.csc_matrix(np.array([[1.0], [-1.0]]))
    r = po.ipm_solve(sp.csc_matrix(np.array([[1.0]])), np.array([0.0]), A, np.array([-1.0, -1.0]), np.array([0, 0], dtype=bool))
    assert r["status"] == 3   # PrimalInfeasible: x <= -1 and x >= 1
    st = po.default_admm_settings()
    r2 = po.admm_solve(sp.csc_matrix(np.array([[1.0]])), np.array([0.0]), sp.csc_matrix(np.array([[1.0], [1.0]])),
                       np.array([-1e30, 1.0]), np.array([-1.0, 1e30]), settings=st)
    assert r2["status"] == 3


def test_random_strictly_convex_qps_satisfy_kkt():
    rng = np.random.default_rng(5)
    for _ in range(10):
        n, me, mi = 12, 3, 20
        M = rng.normal(size=(n, n))
        Pm = M @ M.T + 0.1 * np.eye(n)
        qv = rng.normal(size=n)
        x0 = rng.normal(size=n)
        Ae = rng.normal(size=(me, n))
        Gi = rng.normal(size=(mi, n))
        A = np.vstack([Ae, Gi])
        b = np.concatenate([Ae @ x0, Gi @ x0 + rng.uniform(0.0, 1.0, mi)])
        is_eq = np.arange(me + mi) < me
        r = po.ipm_solve(sp.csc_matrix(Pm), qv, sp.csc_matrix(A), b, is_eq)
        assert r["status"] == 0
        x, y = r["x"], r["y"]
        assert np.abs(Pm @ x + qv + A.T @ y).max() < 1e-6
        res = A @ x - b
        assert np.abs(res[:me]).max() < 1e-7 and res[me:].max() < 1e-7
        assert np.all(y[me:] >= -1e-10) and np.abs(y[me:] * res[me:]).max() < 1e-6
