"""Developer diagnostic: wall time of bgg_targets_from_traj_batch for 4096 robots."""
import sys, time, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from common import wl
from test_oracle_ik import NOMINAL_JOINTS
cfg_name = "a1_configuration"; cfg = wl.CONFIGS[cfg_name]
B = 4096
st, t0, ee = wl.batched_trot_inputs(cfg, B, seed=0)
m = common.make_gpu(cfg_name, B, st); m.SetKinematics(wl.robot())
for _ in range(3): m.GetRealTimeUpdate(st, t0, ee)
q0 = np.concatenate([st[:, :3], st[:, 6:10], np.tile(NOMINAL_JOINTS, (B, 1))], axis=1)
out = m.GetTargetsFromTraj(0.01, q0)
t = time.perf_counter()
for _ in range(5): out = m.GetTargetsFromTraj(0.01, out["q_des"])
dt = (time.perf_counter() - t) / 5
print("GetTargetsFromTraj, 4096 robots (8192 IK solves):", round(1e3 * dt, 2), "ms per call; status hist", np.bincount(out["status"], minlength=4).tolist())
