"""Developer diagnostic: the instances of BASELINE config #2 whose RTI solve does not end `Solved` on the CUDA path,
replayed on the CPU oracle (same inputs, same number of RTI steps) -- do both sides agree on the outcome?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from common import wl
cfg_name = os.environ.get("CFG", "a1_configuration")
cfg = wl.CONFIGS[cfg_name]
B = int(os.environ.get("B", 4096))
STEPS = int(os.environ.get("STEPS", 5))
states, t0, ee = wl.batched_trot_inputs(cfg, B, seed=0)
gpu = common.make_gpu(cfg_name, B, states)
hist = []
for it in range(STEPS):
    out = gpu.GetRealTimeUpdate(states, t0, ee)
    hist.append(out["status"].copy())
    print(it, "status hist", np.bincount(out["status"], minlength=9).tolist(), "iters mean", round(float(out["iters"].mean()), 2))
bad = np.where(~np.isin(hist[-1], (0,)))[0]
print("not Solved at the last step:", len(bad), "of", B)
agree = 0
rows = []
for b in bad[: int(os.environ.get("NBAD", 24))]:
    o = common.make_oracle(cfg_name, states[b])
    seq = []
    for it in range(STEPS):
        seq.append(int(o.solve(states[b], 0.0, ee[b], real_time=True)))
    g = [int(h[b]) for h in hist]
    rows.append((int(b), g, seq))
    agree += int(g[-1] == seq[-1])
for r in rows:
    print("instance", r[0], "cuda", r[1], "oracle", r[2])
print("same final status on", agree, "of", len(rows))
