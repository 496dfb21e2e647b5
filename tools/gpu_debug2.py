"""Developer script: dump one config's first-solve QP (oracle) and both solutions for offline comparison."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from common import wl

cfg_name = os.environ.get("CFG", "a1_gait_opt_config")
cfg = wl.CONFIGS[cfg_name]
B = 6
states, t0, ee = wl.batched_trot_inputs(cfg, B, seed=3)
states[0] = cfg["srb_init"]
ee[0] = wl.EE_NOMINAL
gpu = common.make_gpu(cfg_name, B, states)
out = gpu.GetRealTimeUpdate(states, t0, ee)
dump = {}
for b in range(B):
    o = common.make_oracle(cfg_name, states[b])
    o.assemble(states[b], 0.0, ee[b])
    qp = o.qp()
    o.solve(states[b], 0.0, ee[b], real_time=True)
    oq = o.qp_solution()
    sol = gpu.solution(b)
    sz = gpu.sizes(b)
    obj = lambda z: 0.5 * z @ (qp["P"] @ z) + qp["q"] @ z
    print(b, "gpu status", out["status"][b], "iters", out["iters"][b], "obj", obj(sol["qp_sol"]), "| oracle status", oq["status"], "iters", oq["iters"],
          "obj", obj(oq["x"]), "relerr", np.linalg.norm(sol["qp_sol"] - oq["x"]) / np.linalg.norm(oq["x"]),
          "gpu res", sz["prim_res"], sz["dual_res"], sz["gap"], "oracle res", oq["prim_res"], oq["dual_res"])
    dump[f"b{b}_A"] = qp["A"].toarray(); dump[f"b{b}_P"] = qp["P"].diagonal(); dump[f"b{b}_q"] = qp["q"]; dump[f"b{b}_ub"] = qp["ub"]
    dump[f"b{b}_iseq"] = qp["is_eq"]; dump[f"b{b}_xg"] = sol["qp_sol"]; dump[f"b{b}_xo"] = oq["x"]; dump[f"b{b}_yo"] = oq["dual"]
    dump[f"b{b}_lam"] = sol["lam"]; dump[f"b{b}_slack"] = sol["slack"]; dump[f"b{b}_nueq"] = sol["nu_eq"]
    dump[f"b{b}_sizes"] = np.array([sz[k] for k in ("n", "nu", "nf", "np", "n_samples", "n_eebox", "n_eq", "n_td", "m_ineq")])
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "debug2.npz"), **dump)
