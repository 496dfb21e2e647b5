"""Status histogram of the first RTI solves of BASELINE config #2 (developer diagnostic)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from common import wl
cfg_name = os.environ.get("CFG", "a1_configuration")
cfg = wl.CONFIGS[cfg_name]
B = int(os.environ.get("B", 4096))
states, t0, ee = wl.batched_trot_inputs(cfg, B, seed=0)
gpu = common.make_gpu(cfg_name, B, states)
for it in range(6):
    out = gpu.GetRealTimeUpdate(states, t0, ee)
    print(it, "status hist", np.bincount(out["status"], minlength=9).tolist(), "iters mean", out["iters"].mean(), "max", out["iters"].max(),
          "alpha<1:", int((out["alpha"] < 1).sum()))
bad = np.where(out["status"] == 8)[0][:3]
for b in bad:
    print("Other instance", b, gpu.sizes(b))
