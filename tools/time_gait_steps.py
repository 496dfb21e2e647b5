"""Developer diagnostic: wall time of each piece of a repeated gait-optimisation step (config #3: 64 instances, N = 50, K = 64), step by
step from one snapshot of the instances -- shows where a step's time goes and how much it varies between identical steps."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, common
from common import wl
cfg_name = "a1_gait_opt_config"
cfg = wl.CONFIGS[cfg_name]
B, K = 64, 64
states, t0, ee = wl.batched_trot_inputs(cfg, B, seed=1000)
gpu = common.make_gpu(cfg_name, B, states)
gpu.upload(states, t0, ee)
def one(restore):
    T = []
    def lap(f):
        t = time.perf_counter(); r = f(); gpu.synchronize(); T.append(round(1e3 * (time.perf_counter() - t), 1)); return r
    if restore is not None:
        lap(lambda: [gpu.set_instance(b, restore[b]) for b in range(B)])
    lap(gpu.solve_resident)
    g = lap(gpu.ComputeCostFcnDerivWrtContactTimes)
    lp = lap(lambda: gpu.OptimizeContactTimes(t0))
    ls = lap(lambda: gpu.LineSearch(states, t0, ee, lp["xk"], lp["step"], K=K))
    return T, ls
for _ in range(3):
    one(None)
snap = [gpu.get_instance(b).copy() for b in range(B)]
for i in range(int(os.environ.get("STEPS", 8))):
    T, ls = one(snap)
    print(i, "restore / solve / gradient / LP / line search ms", T, "children status hist", np.bincount(ls["quality"].ravel(), minlength=9).tolist())
