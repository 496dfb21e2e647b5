"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the gait optimiser's derivative path (numpy / scipy on top of
oracle/liboracle.so).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.

Restates, on the oracle's assembled QP and its interior-point solution:
  ClarabelInterface::Computedx / SetupDerivativeCalcs        mpc/qp/clarabel_interface.cpp:604-612, 262-602
  ClarabelInterface::CalcDerivativeWrtVecs / WrtMats         mpc/qp/clarabel_interface.cpp:182-196, 198-260
  MPC::ComputeDerivativeTerms / GetQPPartials                mpc/mpc.cpp:1047-1069
  GaitOptimizer::ModifyQPPartials / ComputeCostFcnDerivWrtContactTimes   mpc/gait_optimizer.cpp:536-539, 92-179
  GaitOptimizer::OptimizeContactTimes and its Create*Constraint helpers  mpc/gait_optimizer.cpp:185-364, 410-534
  GaitOptimizer::GetContactTimes / ConvertQPVecToContactTimes / LineSearch  mpc/gait_optimizer.cpp:645-753
  MPCController::GaitOpt (the call order)                    controllers/mpc_controller.cpp:518-573

The differential system is kept exactly as the reference builds it, including the +diag(slacks) block where the
derivation in its own comment has D(Gz - h) = -diag(slacks):
      [ P    G' D(lam)   A' ]            [ P z + q ]
  d = -[ G    D(s)        0  ]^-1  *      [    0    ]        z = prev_qp_sol (post line search), lam/s from the solver
      [ A    0           0  ]            [    0    ]
Eigen::SparseLU is replaced by scipy's SuperLU (same algorithm family; third-party, unpinned in the reference).
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import pyoracle as po

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
TOUCH_DOWN, LIFT_OFF = po.TOUCH_DOWN, po.LIFT_OFF


def _bind(lib):
    if getattr(lib, "_gait_bound", False):
        return
    lib.orc_mpc_param_partials.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _ip, _ip, _ip, _dp, _ip, _ip, _dp, _dp]
    lib.orc_mpc_num_contacts.argtypes = [C.c_void_p, C.c_int]
    lib.orc_mpc_get_contact_times.argtypes = [C.c_void_p, C.c_int, _dp, _ip]
    lib._gait_bound = True


def contact_times(o):
    """Trajectory::GetContactTimes of the oracle MPC's current trajectory: per foot (times, types)."""
    _bind(o.lib)
    out = []
    for ee in range(4):
        n = o.lib.orc_mpc_num_contacts(o.h, ee)
        t, ty = np.zeros(n), np.zeros(n, np.int32)
        o.lib.orc_mpc_get_contact_times(o.h, ee, t.ctypes.data_as(_dp), ty.ctypes.data_as(_ip))
        out.append((t, ty))
    return out


def param_partials(o, ee, idx, cap=200000):
    """MPCSingleRigidBody::ComputeParamPartialsClarabel for contact time (ee, idx): dict(dA, dG csr; db) or None when the
    last solve was not `Solved`."""
    _bind(o.lib)
    cnt = np.zeros(4, np.int32)
    Ar, Ac, Av = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap)
    Gr, Gc, Gv = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap)
    sz = o.sizes()
    db = np.zeros(sz["num_eq"])
    rc = o.lib.orc_mpc_param_partials(o.h, ee, idx, cap, cnt.ctypes.data_as(_ip), Ar.ctypes.data_as(_ip), Ac.ctypes.data_as(_ip),
                                      Av.ctypes.data_as(_dp), Gr.ctypes.data_as(_ip), Gc.ctypes.data_as(_ip),
                                      Gv.ctypes.data_as(_dp), db.ctypes.data_as(_dp))
    if rc == 1:
        return None
    if rc != 0:
        raise po.OracleError(o.lib.orc_last_error().decode())
    na, ng = int(cnt[0]), int(cnt[1])
    n = sz["n"]
    dA = sp.coo_matrix((Av[:na], (Ar[:na], Ac[:na])), shape=(sz["num_eq"], n)).tocsr()    # setFromTriplets sums duplicates
    dG = sp.coo_matrix((Gv[:ng], (Gr[:ng], Gc[:ng])), shape=(sz["num_ineq"], n)).tocsr()
    return dict(dA=dA, dG=dG, db=db, dh=np.zeros(sz["num_ineq"]), dq=np.zeros(n))


def split_rows(qp):
    """Row index sets of the stacked constraint matrix in the order SetupDerivativeCalcs walks them: inequalities
    (ForceBox | FrictionCone | EndEffectorLocation) and equalities (Dynamics | TDPosition | EndEffectorStart)."""
    eq = np.flatnonzero(qp["is_eq"])
    ineq = np.flatnonzero(~qp["is_eq"])
    return ineq, eq


def derivative_terms(o):
    """MPC::ComputeDerivativeTerms + GetQPPartials.  Returns None unless the last solve was `Solved`."""
    sol = o.qp_solution()
    if sol["status"] != 0:
        return None
    qp = o.qp()
    ineq, eq = split_rows(qp)
    A, P = qp["A"].tocsr(), qp["P"].tocsr()
    G, Ae = A[ineq], A[eq]
    primal, dual, slack = sol["x"], sol["dual"], sol["slack"]
    lam, nu, s = dual[ineq], dual[eq], slack[ineq]
    z = o.prev_qp_sol()
    dx = P @ z + qp["q"]                                                  # Computedx (mpc.cpp:1049)
    n, mi, me = len(z), len(ineq), len(eq)
    # Rows of G without any stored entry (the touch-down sample of every stance: all weights are exactly 0) read
    # 0 + s = h; the oracle's interior point keeps them out of the iteration and reports s = h, which is exactly 0 for
    # the cone / lower-force rows, where Clarabel would return a tiny positive slack.  Either way the row decouples
    # (dlam_i = 0); a unit diagonal keeps the matrix non-singular without changing any other unknown.
    empty = np.diff(G.indptr) == 0
    s = np.where(empty & (s == 0.0), 1.0, s)
    M = sp.bmat([[P, (G.T @ sp.diags(lam)), Ae.T],
                 [G, sp.diags(s), None],
                 [Ae, None, None]], format="csc")
    rhs = np.concatenate([dx, np.zeros(mi + me)])
    d = -spla.splu(M).solve(rhs)
    dz, dlam, dnu = d[:n], d[n:n + mi], d[n + mi:]
    return dict(qp=qp, ineq=ineq, eq=eq, primal=primal, lam=lam, nu=nu, slack=s, z=z, dx=dx, dz=dz, dlam=dlam, dnu=dnu,
                dq=dz.copy(), dh=-lam * dlam, db=-dnu)


def cost_gradient(o, terms=None):
    """dH/dtheta over all contact times, foot-major (GaitOptimizer::ComputeCostFcnDerivWrtContactTimes), after
    ModifyQPPartials(prev_qp_sol).  dA = dnu z*' + nu dz', dG = D(lam) dlam z*' + lam dz' are contracted without being
    formed: <dA, X> = dnu'(X z*) + nu'(X dz)."""
    t = terms or derivative_terms(o)
    if t is None:
        return None
    ct = contact_times(o)
    out = []
    dq = t["dq"] + t["z"]                                                 # ModifyQPPartials
    for ee in range(4):
        for idx in range(len(ct[ee][0])):
            pp = param_partials(o, ee, idx)
            v = t["dnu"] @ (pp["dA"] @ t["primal"]) + t["nu"] @ (pp["dA"] @ t["dz"])
            v += (t["lam"] * t["dlam"]) @ (pp["dG"] @ t["primal"]) + t["lam"] @ (pp["dG"] @ t["dz"])
            v += dq @ pp["dq"] + t["db"] @ pp["db"] + t["dh"] @ pp["dh"]
            out.append(v)
    return np.array(out)


# ---------------------------------------------------------------------------------------------- contact-time LP
def gait_lp(ct, grad, time, trust=1.0, min_time=0.2):
    """The LP of GaitOptimizer::OptimizeContactTimes (P = Bk = 0): rows = polytope | start | trust region | next-node,
    num_constraints = 2 n + 3 * num_ee (unused trailing rows stay 0 <= 0 x <= 0).  Returns (A csr, lb, ub)."""
    counts = [len(t) for t, _ in ct]
    base = np.concatenate([[0], np.cumsum(counts)])
    n = int(base[-1])
    m = 2 * n + 3 * 4
    A = sp.lil_matrix((m, n))
    lb, ub = np.zeros(m), np.zeros(m)
    row = 0
    for ee in range(4):                                                   # CreatePolytopeConstraint, :410-465
        t, ty = ct[ee]
        nodes = len(t)
        next_node = -1
        for j in range(1, nodes):
            if t[j] >= time:
                next_node = j
                break
        if ty[next_node] == TOUCH_DOWN:
            ub[row + base[ee] + next_node - 1] = t[next_node] - t[next_node - 1]
            lb[row + base[ee] + next_node - 1] = -3
        for i in range(1, nodes):
            A[row + base[ee] + i - 1, base[ee] + i - 1] = 1
            A[row + base[ee] + i - 1, base[ee] + i] = -1
            if i != next_node or ty[next_node] != TOUCH_DOWN:
                ub[row + base[ee] + i - 1] = t[i] - t[i - 1] - min_time
                lb[row + base[ee] + i - 1] = -2
        A[row + base[ee] + nodes - 1, base[ee] + nodes - 1] = 1
        lb[row + base[ee] + nodes - 1] = 0
        ub[row + base[ee] + nodes - 1] = 1
    row = n
    for ee in range(4):                                                   # CreateStartConstraint, :492-500
        A[row + ee, base[ee]] = 1
    row += 4
    for i in range(n):                                                    # CreateTrustRegionConstraint, :502-511
        A[row + i, i] = 1
        lb[row + i], ub[row + i] = -trust, trust
    row += n
    k = 0
    for ee in range(4):                                                   # CreateNextNodeConstraints, :513-534
        t, ty = ct[ee]
        next_node = -1
        for i in range(1, len(t)):
            if t[i] >= time:
                next_node = i
                break
        if ty[next_node] == TOUCH_DOWN:
            A[row + k, base[ee] + next_node - 1] = 1
            A[row + k + 1, base[ee] + next_node] = 1
            k += 2
    return A.tocsr(), lb, ub


def solve_gait_lp(ct, grad, time, trust=1.0):
    """Optimum of the LP (the reference runs OSQP with eps 1e-10 and polishing, i.e. to a vertex; scipy's HiGHS simplex
    gives the same vertex when the optimum is unique).  Returns the step."""
    from scipy.optimize import linprog
    A, lb, ub = gait_lp(ct, grad, time, trust)
    Ad = A.toarray()
    res = linprog(grad, A_ub=np.vstack([Ad, -Ad]), b_ub=np.concatenate([ub, -lb]), bounds=(None, None), method="highs-ds")
    if res.status != 0:
        raise po.OracleError("gait LP: " + res.message)
    return res.x


def contact_times_for(ct, xk, step, alpha):
    """GaitOptimizer::GetContactTimes(alpha) -> ConvertQPVecToContactTimes, :645-669."""
    vec = xk + alpha * step
    out, k = [], 0
    for ee in range(4):
        t = ct[ee][0].copy()
        for i in range(len(t)):
            t[i] = vec[k + i]
            if i > 0 and 0 < t[i - 1] - t[i] <= 1e-3:
                t[i] = t[i - 1]
        k += len(t)
        out.append(t)
    return out


def line_search(o, state, time, ee_locations, ct, xk, step, ls_size=10):
    """GaitOptimizer::LineSearch, :671-753: one RTI solve per alpha_i = i / LS_SIZE on a copy of the MPC; argmin of
    cost / num_decision_vars over the copies that are not primal infeasible."""
    costs, quality = [], []
    for i in range(ls_size):
        c = o.clone()
        times = contact_times_for(ct, xk, step, i / ls_size)
        for ee in range(4):
            c.set_contact_times(ee, times[ee])
        c.solve(state, time, ee_locations, real_time=True)
        costs.append(c.cost() / c.sizes()["n"])
        quality.append(int(c.qp_solution()["status"]))
    best, cmin = -1, 1e10
    for i in range(ls_size):
        if costs[i] < cmin and quality[i] != 3:
            cmin, best = costs[i], i
    return best, np.array(costs), np.array(quality)
