"""Small fixed workload for ncu: the gait-optimisation tick (solve, gait gradient, contact-time LP, line search) of BASELINE
config #3 -- B instances of a1_gait_opt_config (N = 50), K candidates."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bilevel-gait-gen_b200"))
import bgg_b200 as bg   # noqa: E402
import workloads as wl  # noqa: E402

B, K = int(os.environ.get("B", 64)), int(os.environ.get("K", 10))
cfg_name = os.environ.get("CFG", "a1_gait_opt_config")
cfg = wl.CONFIGS[cfg_name]
states = np.tile(np.asarray(cfg["srb_init"], float), (B, 1))
states[:, :2] += np.random.default_rng(0).uniform(-0.01, 0.01, (B, 2))
t0, ee = np.zeros(B), np.tile(wl.EE_NOMINAL, (B, 1, 1))
mpc = bg.BatchedMPC(cfg["num_nodes"], cfg["integrator_dt"], wl.robot(), **wl.mpc_kwargs(cfg))
mpc.AddQuadraticTrackingCost(wl.target_tangent(cfg), np.asarray(cfg["Q"], float))
mpc.Reset(B)
mpc.SetStateTrajectoryWarmStart(states)
for _ in range(3):
    out = mpc.GetRealTimeUpdate(states, t0, ee)
g = mpc.ComputeCostFcnDerivWrtContactTimes()
lp = mpc.OptimizeContactTimes(t0)
ls = mpc.LineSearch(states, t0, ee, lp["xk"], lp["step"], K=K)
print("status", np.bincount(out["status"], minlength=9).tolist(), "gradients", int((g["status"] == 0).sum()), "best", np.bincount(np.maximum(ls["best"], 0), minlength=K).tolist())
