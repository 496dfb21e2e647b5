"""Developer diagnostic: status histograms of consecutive RTI solves of BASELINE config #2 for two builds of the library
(LIB_A / LIB_B environment variables; default: the in-tree build)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from common import wl
import bgg_b200 as bg
cfg_name = "a1_configuration"
cfg = wl.CONFIGS[cfg_name]
B = int(os.environ.get("B", 4096))
states, t0, ee = wl.batched_trot_inputs(cfg, B, seed=0)
res = {}
for tag, lib in (("A", os.environ.get("LIB_A")), ("B", os.environ.get("LIB_B"))):
    if not lib:
        continue
    bg.LIB_PATH = lib
    gpu = common.make_gpu(cfg_name, B, states)
    hist = []
    for it in range(5):
        out = gpu.GetRealTimeUpdate(states, t0, ee)
        hist.append(out["status"].copy())
        print(tag, os.path.basename(os.path.dirname(lib)), it, np.bincount(out["status"], minlength=9).tolist(), "iters", round(float(out["iters"].mean()), 2))
    res[tag] = hist
if len(res) == 2:
    for it in range(5):
        a, b = res["A"][it], res["B"][it]
        print("step", it, "A solved & B not:", int(((a == 0) & (b != 0)).sum()), " B solved & A not:", int(((b == 0) & (a != 0)).sum()))
