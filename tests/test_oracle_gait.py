"""CPU tests of the oracle's gait-optimiser derivative path (oracle/gait_partials.cpp, oracle/gait_oracle.py).

The reference pins this path with finite differences (test/mpc_test.cpp:114-236: dynamics, force-box and friction-cone
blocks of the assembled constraint matrix against ComputeParamPartialsClarabel, step sqrt(1e-16), margin 1e-4); the same
check is restated here against the oracle's own assembly.  MPC-scale values of the adjoint itself have no stored vectors
in the reference ("parity unpinned"); they are checked through the linear system they must satisfy."""
import numpy as np
import pytest

import common
from common import wl

go = pytest.importorskip("gait_oracle")


def _solved_oracle(cfg_name="a1_configuration"):
    cfg = wl.CONFIGS[cfg_name]
    o = common.make_oracle(cfg_name)
    init = np.asarray(cfg["srb_init"], float)
    ee = wl.EE_NOMINAL.copy()
    o.initial_run(init, ee)
    assert o.qp_solution()["status"] == 0
    return o, init, ee


def test_param_partials_match_finite_differences_of_the_assembly():
    o, init, ee = _solved_oracle()
    sz = o.sizes()
    base = o.clone()
    base.assemble(init, 0.0, ee)
    A1 = base.qp()["A"].toarray()
    nd, nfb, nc = sz["num_dyn"], sz["num_force_box"], sz["num_cone"]
    h, margin = 1e-8, 1e-4
    ct = go.contact_times(o)
    checked = 0
    for foot in range(4):
        for idx in range(1, len(ct[foot][0])):
            t = ct[foot][0].copy()
            t[idx] += h
            o2 = o.clone()
            o2.set_contact_times(foot, t)
            o2.assemble(init, 0.0, ee)
            A2 = o2.qp()["A"].toarray()
            assert A2.shape == A1.shape
            fd = (A2 - A1) / h
            pp = go.param_partials(o, foot, idx)
            dA, dG = pp["dA"].toarray(), pp["dG"].toarray()
            assert np.abs(dA[:nd] - fd[:nd]).max() < margin, (foot, idx, "dynamics")
            assert np.abs(dG[:nfb] - fd[nd:nd + nfb]).max() < margin, (foot, idx, "force box")
            assert np.abs(dG[nfb:nfb + nc] - fd[nd + nfb:nd + nfb + nc]).max() < margin, (foot, idx, "cone")
            checked += 1
    assert checked == 16


def test_adjoint_solves_the_reference_differential_system():
    o, _, _ = _solved_oracle()
    t = go.derivative_terms(o)
    qp = t["qp"]
    A, P = qp["A"].tocsr(), qp["P"]
    G, Ae = A[t["ineq"]], A[t["eq"]]
    lam, s, dz, dlam, dnu = t["lam"], t["slack"], t["dz"], t["dlam"], t["dnu"]
    r1 = P @ dz + G.T @ (lam * dlam) + Ae.T @ dnu + t["dx"]
    empty = np.diff(G.indptr) == 0
    r2 = (G @ dz + s * dlam)[~empty]
    r3 = Ae @ dz
    scale = max(1.0, np.abs(t["dx"]).max())
    assert np.abs(r1).max() < 1e-6 * scale and np.abs(r2).max() < 1e-8 and np.abs(r3).max() < 1e-8
    assert np.all(dlam[empty] == 0.0)
    assert np.array_equal(t["dq"], dz) and np.array_equal(t["db"], -dnu) and np.array_equal(t["dh"], -lam * dlam)


def test_gradient_and_lp_step():
    o, init, ee = _solved_oracle()
    g = go.cost_gradient(o)
    ct = go.contact_times(o)
    assert g.shape == (sum(len(t) for t, _ in ct),) and np.all(np.isfinite(g))
    # the contraction without forming dA / dG equals the dense definition (gait_optimizer.cpp:92-179)
    t = go.derivative_terms(o)
    pp = go.param_partials(o, 1, 2)
    dA = np.outer(t["dnu"], t["primal"]) + np.outer(t["nu"], t["dz"])
    dG = np.outer(t["lam"] * t["dlam"], t["primal"]) + np.outer(t["lam"], t["dz"])
    dense = (dA * pp["dA"].toarray()).sum() + (dG * pp["dG"].toarray()).sum() + t["db"] @ pp["db"]
    assert abs(dense - g[5 + 2]) <= 1e-9 * max(1.0, abs(dense))
    A, lb, ub = go.gait_lp(ct, g, 0.0)
    assert A.shape == (2 * 20 + 12, 20)
    step = go.solve_gait_lp(ct, g, 0.0)
    r = A @ step
    assert np.all(r <= ub + 1e-9) and np.all(r >= lb - 1e-9)
    assert g @ step <= 1e-12                      # a descent step of the linear model
    for foot in range(4):                          # first contact time never moves (CreateStartConstraint)
        assert abs(step[5 * foot]) < 1e-12
    best, costs, quality = go.line_search(o, init, 0.0, ee, ct, np.concatenate([t for t, _ in ct]), step, ls_size=4)
    assert 0 <= best < 4 and np.all(np.isfinite(costs))
