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
    nd = 12 * (N + 1); all_terms = []
    for b in range(B):
        o = oracles[b]; terms = go.derivative_terms(o); all_terms.append(terms)
        sol = gpu.solution(b); order = common.gpu_rows_to_reference_order(sol, N)
        lam_k, s_k = np.zeros_like(sol["lam"]), np.zeros_like(sol["slack"])
        lam_k[order], s_k[order] = terms["lam"], terms["slack"]
        common.mirror_oracle_to_gpu(o, gpu, b)
        gpu.set_solution(b, qp_sol=terms["primal"], z=terms["z"], lam=lam_k, slack=s_k, nu_eq=terms["nu"][nd:])
    res = gpu.ComputeCostFcnDerivWrtContactTimes()
    for b in range(B):
        terms = all_terms[b]; adj = gpu.adjoint(b)
        g_o = go.cost_gradient(oracles[b], terms)
        print(cfg_name, b, "status", res["status"][b], "rel dz", T._rel(adj["dz"], terms["dz"]), "abs err", np.abs(adj["dz"] - terms["dz"]).max(), "|dz|", np.abs(terms["dz"]).max(),
              "|dlam|", np.abs(terms["dlam"]).max(), "|dnu|", np.abs(terms["dnu"]).max(), "rel dnu_dyn", T._rel(adj["dnu_dyn"], terms["dnu"][:nd]), "grad", np.abs(res["dHdtheta"][b] - g_o).max() / max(1.0, np.abs(g_o).max()))
