import sys, os
sys.path.insert(0, "tests")
import numpy as np, common
from common import wl
for cfg_name in ["a1_configuration", "a1_gait_opt_config"]:
    cfg = wl.CONFIGS[cfg_name]
    B = 12
    states, t0, ee = wl.batched_trot_inputs(cfg, B, seed=3)
    for tol in (1e-8, 1e-9, 1e-10):
        gpu = common.make_gpu(cfg_name, B, states, ipm_tol=tol, ipm_tol_gap=tol)
        out = gpu.GetRealTimeUpdate(states, t0, ee)
        rels = []; its = []
        for b in range(B):
            o = common.make_oracle(cfg_name, states[b]); o.set_ipm(tol_feas=tol, tol_gap=tol)
            st = o.solve(states[b], 0.0, ee[b], real_time=True)
            oq = o.qp_solution(); sol = gpu.solution(b)
            rels.append(np.linalg.norm(sol["qp_sol"] - oq["x"]) / np.linalg.norm(oq["x"])); its.append((int(out["iters"][b]), oq["iters"], int(out["status"][b]), st))
        print(cfg_name, tol, "max rel", max(rels), "median", np.median(rels), its)
