"""Debug helper: gait gradient, CUDA path vs oracle, verbose."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "bilevel-gait-gen_b200")]
import common
from common import wl
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity as T

cfg_name = os.environ.get("CFG", "a1_configuration")
cfg = wl.CONFIGS[cfg_name]
N = cfg["num_nodes"]
B = 3
states, _, ee = wl.batched_trot_inputs(cfg, B, seed=21)
states[0] = cfg["srb_init"]; ee[0] = wl.EE_NOMINAL
gpu, oracles, out, go = T._gradient_case(cfg_name, states, ee, tol_gap=float(os.environ.get('GAP', 0)))
res = gpu.ComputeCostFcnDerivWrtContactTimes()
print("iters", out["iters"], [o.qp_solution()["iters"] for o in oracles]); print("solve status", out["status"], "grad status", res["status"], res["n_contacts"])
rel = lambda a, b: np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
for b in range(B):
    o = oracles[b]
    terms = go.derivative_terms(o)
    if terms is None:
        print(b, "oracle not solved"); continue
    adj = gpu.adjoint(b); sol = gpu.solution(b)
    order = common.gpu_rows_to_reference_order(sol, N)
    nd = 12 * (N + 1)
    print(b, "qp_sol rel", rel(sol["qp_sol"], terms["primal"]), "z rel", rel(sol["z"], terms["z"]))
    print("  lam rel", rel(sol["lam"][order], terms["lam"]), "slack rel", rel(sol["slack"][order], terms["slack"]))
    print("  dz", rel(adj["dz"], terms["dz"]), "nu", rel(adj["nu_dyn"], terms["nu"][:nd]), "dnu", rel(adj["dnu_dyn"], terms["dnu"][:nd]),
          "dnu_eq", rel(adj["dnu_eq"], terms["dnu"][nd:]), np.abs(terms["dnu"][nd:]).max())
    y_g, y_o = adj["dlam"][order] * sol["lam"][order], terms["dlam"] * terms["lam"]
    print("  lam*dlam max diff", np.abs(y_g - y_o).max(), "scale", np.abs(y_o).max())
    g_o = go.cost_gradient(o, terms); g = res["dHdtheta"][b]
    print("  grad gpu", np.array2string(g, precision=4))
    print("  grad orc", np.array2string(g_o, precision=4))
    print("  max abs diff", np.abs(g - g_o).max(), "scale", np.abs(g_o).max())

print("---- injected oracle solution ----")
for b in range(B):
    o = oracles[b]
    terms = go.derivative_terms(o)
    if terms is None:
        continue
    sol = gpu.solution(b)
    order = common.gpu_rows_to_reference_order(sol, N)
    lam_k, s_k = np.zeros_like(sol["lam"]), np.zeros_like(sol["slack"])
    lam_k[order], s_k[order] = terms["lam"], terms["slack"]
    nd = 12 * (N + 1)
    common.mirror_oracle_to_gpu(o, gpu, b)
    gpu.set_solution(b, qp_sol=terms["primal"], z=terms["z"], lam=lam_k, slack=s_k, nu_eq=terms["nu"][nd:])
res = gpu.ComputeCostFcnDerivWrtContactTimes()
for b in range(B):
    o = oracles[b]
    terms = go.derivative_terms(o)
    if terms is None:
        continue
    adj = gpu.adjoint(b)
    nd = 12 * (N + 1)
    print(b, "dz", rel(adj["dz"], terms["dz"]), "|dz|", np.linalg.norm(terms["dz"]), "nu", rel(adj["nu_dyn"], terms["nu"][:nd]), "dnu", rel(adj["dnu_dyn"], terms["dnu"][:nd]),
          "dnu_eq", rel(adj["dnu_eq"], terms["dnu"][nd:]))
    g_o = go.cost_gradient(o, terms); g = res["dHdtheta"][b]
    print("  max abs diff", np.abs(g - g_o).max(), "scale", np.abs(g_o).max())
