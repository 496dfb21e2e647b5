"""Developer script (runs on the GPU box via gpurun): first-contact check of the CUDA path against the oracle.
Dumps every intermediate to gpurun_out/debug1.npz so the comparison can be redone offline."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "bilevel-gait-gen_b200"))
import bgg_b200 as bg   # noqa: E402
import pyoracle as po   # noqa: E402

po.build()
rc = po.load_robot_consts(os.path.join(ROOT, "tests", "golden", "a1_robot_consts.json"))
N, dt = int(os.environ.get("N", 20)), float(os.environ.get("DT", 0.05))
Q = np.array([340, 340, 4000, 0.1, 0.1, 10, 3000, 3000, 3000, 1, 1, 1.0])
init = np.array([0, 0, 0.3, 0, 0, 0, 0, 0, 0, 1.0, 0, 0, 0])
des = np.array([0, 0, 0.3, 0, 0, 0, 0, 0, 0, 0, 0, 0.0])
ee = np.array([[0.1526, 0.12523, 0.011089], [0.1526, -0.12523, 0.011089], [-0.208321844, 0.1363286, 0.01444],
               [-0.208321844, -0.1363286, 0.01444]])
B = 4
rng = np.random.default_rng(0)
states = np.tile(init, (B, 1))
for b in range(1, B):
    states[b, :3] += rng.uniform(-1, 1, 3) * [0.05, 0.05, 0.02]
    states[b, 3:6] += rng.uniform(-1, 1, 3) * [1.0, 1.0, 0.5]
    aa = rng.uniform(-1, 1, 3) * 0.1
    states[b, 6:10] = po.quat_exp3(aa)
    states[b, 10:] += rng.uniform(-1, 1, 3) * 0.05
ees = np.tile(ee, (B, 1, 1))
ees[1:, :, :2] += rng.uniform(-0.02, 0.02, (B - 1, 4, 2))

mpc = bg.BatchedMPC(N, dt, rc)
mpc.AddQuadraticTrackingCost(des, Q)
mpc.Reset(B)
mpc.SetStateTrajectoryWarmStart(states)
oracles = []
for b in range(B):
    o = po.SrbMpc(N, dt, rc)
    o.set_costs(des, Q)
    o.set_warm_states(np.tile(states[b], (N + 1, 1)))
    oracles.append(o)

dump = {}
nsolves = int(os.environ.get("NSOLVES", 3))
for it in range(nsolves):
    t = time.time()
    out = mpc.GetRealTimeUpdate(states, 0.0, ees)
    el = time.time() - t
    print(f"--- solve {it}: gpu status {out['status']} iters {out['iters']} alpha {out['alpha']} cost {out['cost']} ({el*1e3:.1f} ms)")
    for b in range(B):
        o = oracles[b]
        sz = mpc.sizes(b)
        o.assemble(states[b], 0.0, ees[b])
        osz = o.sizes()
        Ad, Bd, cd = mpc.dynamics(b, 1)
        oAd, oBd, ocd = o.node_dynamics()
        eA, eB, ec = np.abs(Ad[0] - oAd).max(), np.abs(Bd[0] - oBd).max(), np.abs(cd[0] - ocd).max()
        pat = ((Ad[0] != 0) == (oAd != 0)).all() and ((Bd[0] != 0) == (oBd != 0)).all()
        print(f" b={b} n {sz['n']}/{osz['n']} nu {sz['nu']} m_ineq {sz['m_ineq']}/{osz['num_ineq']} n_eq {sz['n_eq']} err {sz['error']}"
              f" | dyn err A {eA:.2e} B {eB:.2e} c {ec:.2e} pattern_equal {pat}")
        cdn = mpc.condensed(b)
        sol = mpc.solution(b)
        qp = o.qp()
        dump[f"s{it}_b{b}_H"] = cdn["H"]; dump[f"s{it}_b{b}_g"] = cdn["g"]; dump[f"s{it}_b{b}_phipos"] = cdn["phipos"]
        dump[f"s{it}_b{b}_xoff"] = cdn["xoff"]; dump[f"s{it}_b{b}_qpsol"] = sol["qp_sol"]; dump[f"s{it}_b{b}_z"] = sol["z"]
        dump[f"s{it}_b{b}_lam"] = sol["lam"]; dump[f"s{it}_b{b}_slack"] = sol["slack"]; dump[f"s{it}_b{b}_nueq"] = sol["nu_eq"]
        dump[f"s{it}_b{b}_A"] = qp["A"].toarray(); dump[f"s{it}_b{b}_P"] = qp["P"].diagonal(); dump[f"s{it}_b{b}_q"] = qp["q"]
        dump[f"s{it}_b{b}_ub"] = qp["ub"]; dump[f"s{it}_b{b}_iseq"] = qp["is_eq"]
        dump[f"s{it}_b{b}_Ad"] = Ad[0]; dump[f"s{it}_b{b}_Bd"] = Bd[0]; dump[f"s{it}_b{b}_cd"] = cd[0]
        dump[f"s{it}_b{b}_sizes"] = np.array([sz[k] for k in ("n", "nu", "nf", "np", "n_samples", "n_eebox", "n_eq", "n_td", "m_ineq", "status", "iters")])
        # KKT check of the GPU's QP optimum against the oracle's sparse QP
        A, P, q, ub, iseq = qp["A"], qp["P"], qp["q"], qp["ub"], qp["is_eq"]
        z = sol["qp_sol"]
        r = A @ z - ub
        print(f"      sparse-QP check: eq viol {np.abs(r[iseq]).max():.2e} ineq viol {np.maximum(r[~iseq], 0).max():.2e} "
              f"obj {0.5 * z @ (P @ z) + q @ z:.6f} status {bg.STATUS_NAMES[sz['status']]} iters {sz['iters']} "
              f"pres {sz['prim_res']:.1e} dres {sz['dual_res']:.1e} gap {sz['gap']:.1e}")
        # oracle: full solve of the same step for the line-search / trajectory comparison
        o.solve(states[b], 0.0, ees[b], real_time=True)
        ost = o.stats()
        oq = o.qp_solution()
        zo = o.prev_qp_sol()
        print(f"      qp_sol relerr {np.linalg.norm(oq['x'] - z) / np.linalg.norm(oq['x']):.2e} (oracle ipm iters {oq['iters']} status {oq['status']})"
              f" z relerr {np.linalg.norm(zo - sol['z']) / np.linalg.norm(zo):.2e} states err {np.abs(o.states() - mpc.GetStates(b)).max():.2e}"
              f" qp obj gpu {sz['qp_cost']:.8f} oracle {0.5 * oq['x'] @ (P @ oq['x']) + q @ oq['x']:.8f}")
        print(f"      oracle(IPM) cost {ost['cost']:.6f} alpha {ost['alpha']} eqviol {ost['eq_violation']:.3e} | gpu cost {sz['cost']:.6f} "
              f"alpha {sz['alpha']} eqviol {sz['eq_violation']:.3e} merit {sz['merit']:.4f}/{ost['merit']:.4f}")
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "debug1.npz"), **dump)
print("dumped")
