import sys, os
sys.path.insert(0, "tests")
import numpy as np, common
from common import wl
import test_gpu_parity as T
for cfg_name in ["a1_configuration", "a1_gait_opt_config"]:
    cfg = wl.CONFIGS[cfg_name]; N = cfg["num_nodes"]; B = 3
    states, _, ee = wl.batched_trot_inputs(cfg, B, seed=21)
    states[0] = cfg["srb_init"]; ee[0] = wl.EE_NOMINAL
    gpu, oracles, out, go = T._gradient_case(cfg_name, states, ee)
    res = gpu.ComputeCostFcnDerivWrtContactTimes()
    nd = 12 * (N + 1)
    for b in range(B):
        o = oracles[b]; terms = go.derivative_terms(o); sol = gpu.solution(b); adj = gpu.adjoint(b)
        order = common.gpu_rows_to_reference_order(sol, N)
        qp = o.qp(); A = qp["A"]; ine = ~qp["is_eq"]
        Ain = A[np.flatnonzero(ine)]
        lam_g, lam_o = sol["lam"][order], terms["lam"]
        print(cfg_name, b, "iters", int(out["iters"][b]), o.qp_solution()["iters"], "rel lam", T._rel(lam_g, lam_o), "rel A'lam", T._rel(Ain.T @ lam_g, Ain.T @ lam_o),
              "strict compl min(lam+s)", float(np.min(terms["lam"] + terms["slack"])), "rel primal", T._rel(sol["qp_sol"], terms["primal"]),
              "rel dz", T._rel(adj["dz"], terms["dz"]), "|dz|", np.abs(terms["dz"]).max(), "grad rel", np.abs(res["dHdtheta"][b] - go.cost_gradient(o, terms)).max() / max(1, np.abs(go.cost_gradient(o, terms)).max()))
