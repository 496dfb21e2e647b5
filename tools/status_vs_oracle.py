"""Developer diagnostic: BASELINE config #2 (4096 random instances, 5 RTI steps) on the CUDA path, and a sample of the
instances -- every one that is not `Solved` at any step plus the first NSAMPLE -- replayed on the CPU oracle
(same inputs, same number of RTI steps): status and iteration count per step on both sides."""
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
NS = int(os.environ.get("NSAMPLE", 64))
states, t0, ee = wl.batched_trot_inputs(cfg, B, seed=int(os.environ.get("SEED", 0)))
kw = {}
if os.environ.get("REFINE"): kw["ipm_refine"] = int(os.environ["REFINE"])
if os.environ.get("REFINE_AFTER"): kw["ipm_refine_after"] = int(os.environ["REFINE_AFTER"])
gpu = common.make_gpu(cfg_name, B, states, **kw)
hist, ith = [], []
import time
for it in range(STEPS):
    tic = time.perf_counter()
    out = gpu.GetRealTimeUpdate(states, t0, ee)
    print("   wall ms", round(1e3 * (time.perf_counter() - tic), 2), "refined iters mean", np.mean([gpu.sizes(b)["refined_iters"] for b in range(0, B, max(1, B // 64))]))
    hist.append(out["status"].copy()); ith.append(out["iters"].copy())
    print(it, "status hist", np.bincount(out["status"], minlength=9).tolist(), "iters mean", round(float(out["iters"].mean()), 2),
          "max", int(out["iters"].max()), flush=True)
hist = np.array(hist); ith = np.array(ith)
bad = np.where((hist != 0).any(0))[0]
sample = sorted(set(bad.tolist()) | set(range(NS)))
print("not Solved at some step:", len(bad), "of", B, "; oracle sample", len(sample))


def replay(b):
    o = common.make_oracle(cfg_name, states[b])
    st, its = [], []
    for it in range(STEPS):
        st.append(int(o.solve(states[b], 0.0, ee[b], real_time=True)))
        its.append(int(o.qp_solution()["iters"]))
    return b, st, its


import multiprocessing as mp
with mp.Pool(min(16, os.cpu_count())) as pool:
    rows = pool.map(replay, sample)
same = 0; dit = []
for b, st, its in rows:
    g = hist[:, b].tolist()
    same += int(g == st)
    dit += (ith[:, b] - np.array(its)).tolist()
    if g != st or b in bad:
        print("instance", b, "cuda", g, ith[:, b].tolist(), "oracle", st, its)
print("same status sequence on", same, "of", len(rows), "; iteration-count difference histogram", dict(zip(*np.unique(dit, return_counts=True))))
