import sys, os
sys.path.insert(0, "tests")
import numpy as np
import common
from common import wl
cfg_name = "a1_configuration"
cfg = wl.CONFIGS[cfg_name]
B = 8
states, t0, ee = wl.batched_trot_inputs(cfg, B, seed=1)
ox = []
for b in range(B):
    o = common.make_oracle(cfg_name, states[b])
    st = o.solve(states[b], 0.0, ee[b], real_time=True)
    ox.append((st, o.qp_solution()["x"].copy(), o.qp()))
for ra in (-1, 0, 6):
    gpu = common.make_gpu(cfg_name, B, states, ipm_refine_after=ra)
    out = gpu.GetRealTimeUpdate(states, t0, ee)
    rels = []
    for b in range(B):
        st, x, qp = ox[b]
        sol = gpu.solution(b)["qp_sol"]
        obj = lambda z: 0.5 * z @ (qp["P"] @ z) + qp["q"] @ z
        rels.append((st, int(out["status"][b]), float(np.linalg.norm(sol - x) / np.linalg.norm(x)), float((obj(sol) - obj(x)) / max(1, abs(obj(x))))))
    print("refine_after", ra, [(a, b, f"{c:.1e}", f"{d:.1e}") for a, b, c, d in rels], out["iters"].tolist())
# who is off on instance 0: oracle at 1e-8, CUDA at 1e-8, both against the oracle at 1e-11
b = 0
o = common.make_oracle(cfg_name, states[b])
o.set_ipm(tol_feas=1e-11, tol_gap=1e-11, max_iter=200)
st = o.solve(states[b], 0.0, ee[b], real_time=True)
xt = o.qp_solution()["x"].copy()
print("tight oracle status", st, "iters", o.qp_solution()["iters"])
print("oracle(1e-8) vs tight", np.linalg.norm(ox[b][1] - xt) / np.linalg.norm(xt))
for tol in (0.0, 1e-10):
    gpu = common.make_gpu(cfg_name, B, states, ipm_tol=tol, ipm_tol_gap=tol)
    out = gpu.GetRealTimeUpdate(states, t0, ee)
    print("cuda tol", tol, "status", out["status"][b], "iters", out["iters"][b], "vs tight", np.linalg.norm(gpu.solution(b)["qp_sol"] - xt) / np.linalg.norm(xt))
