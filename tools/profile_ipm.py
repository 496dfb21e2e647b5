"""Small fixed workload for ncu: B instances of BASELINE config #2, one warm-up solve then one profiled solve."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bilevel-gait-gen_b200"))
import bgg_b200 as bg   # noqa: E402
import workloads as wl  # noqa: E402

B = int(os.environ.get("B", 296))
cfg_name = os.environ.get("CFG", "a1_configuration")
cfg = wl.CONFIGS[cfg_name]
states, t0, ee = wl.batched_trot_inputs(cfg, B, seed=0)
mpc = bg.BatchedMPC(cfg["num_nodes"], cfg["integrator_dt"], wl.robot(), **wl.mpc_kwargs(cfg))
mpc.AddQuadraticTrackingCost(wl.target_tangent(cfg), np.asarray(cfg["Q"], float))
mpc.Reset(B)
mpc.SetStateTrajectoryWarmStart(states)
for _ in range(int(os.environ.get("SOLVES", 2))):
    out = mpc.GetRealTimeUpdate(states, t0, ee)
print("status counts", np.bincount(out["status"], minlength=9).tolist(), "mean iters", out["iters"].mean())
