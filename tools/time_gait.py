import sys, time
sys.path.insert(0, "tests")
import numpy as np, common
from common import wl
cfg_name = "a1_gait_opt_config"
cfg = wl.CONFIGS[cfg_name]
B = 64
states, t0, ee = wl.batched_trot_inputs(cfg, B, seed=0)
gpu = common.make_gpu(cfg_name, B, states)
for _ in range(3):
    out = gpu.GetRealTimeUpdate(states, t0, ee)
def tm(f, n=5):
    f(); gpu.synchronize() if hasattr(gpu, "synchronize") else None
    t = time.perf_counter()
    for _ in range(n): r = f()
    return (time.perf_counter() - t) / n * 1e3, r
ms, out = tm(lambda: gpu.GetRealTimeUpdate(states, t0, ee)); print("solve 64:", round(ms, 2), "ms")
ms, g = tm(lambda: gpu.ComputeCostFcnDerivWrtContactTimes()); print("gradient 64:", round(ms, 2), "ms", "ok", int((g["status"] == 0).sum()))
ms, lp = tm(lambda: gpu.OptimizeContactTimes(t0)); print("LP 64:", round(ms, 2), "ms")
ms, ls = tm(lambda: gpu.LineSearch(states, t0, ee, lp["xk"], lp["step"], K=64), n=3); print("line search 64x64:", round(ms, 2), "ms")
# iteration / status distribution of the line-search children
import ctypes
hist = np.bincount(ls["quality"].ravel(), minlength=9)
print("children status hist", hist.tolist())
