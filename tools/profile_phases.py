"""Per-phase cycle counts of k_ipm (CTA 0, thread 0) from the -DBGG_IPM_PROF build (tools/build_prof.sh).
B = 1 shows the latency of a CTA alone on its SM, B = 296 / 4096 the same under 2 CTAs per SM."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bilevel-gait-gen_b200"))
import bgg_b200 as bg   # noqa: E402
bg.LIB_PATH = os.path.join(ROOT, "bilevel-gait-gen_b200", "libbgg_b200_prof.so")
import workloads as wl  # noqa: E402

NAMES = {0: "setup", 1: "kkt_assemble", 2: "chol panel solve", 3: "chol trailing+lookahead (to barrier)", 4: "chol invert diag blocks",
         5: "chol_solve", 6: "apply_C", 7: "add_Ct", 8: "apply_H", 9: "apply_E/add_Et", 10: "vector ops / reductions",
         11: "chol first diag block", 12: "(warp 0 own time inside phase 3)", 13: "block reductions", 14: "pass set-up (rows a2, W a2; x = rhs)",
         15: "total direction + step ratios"}
cfg_name = os.environ.get("CFG", "a1_configuration")
cfg = wl.CONFIGS[cfg_name]
for B in [int(x) for x in os.environ.get("BS", "1,296,4096").split(",")]:
    states, t0, ee = wl.batched_trot_inputs(cfg, B, seed=0)
    mpc = bg.BatchedMPC(cfg["num_nodes"], cfg["integrator_dt"], wl.robot(), **wl.mpc_kwargs(cfg))
    mpc.AddQuadraticTrackingCost(wl.target_tangent(cfg), np.asarray(cfg["Q"], float))
    mpc.Reset(B)
    mpc.SetStateTrajectoryWarmStart(states)
    for _ in range(2):
        out = mpc.GetRealTimeUpdate(states, t0, ee)
    prof = (C.c_longlong * 32)()
    lib = C.CDLL(bg.LIB_PATH)
    assert lib.bgg_debug_ipm_prof(prof) == 0
    p = np.array(prof[:], dtype=np.int64)
    it = int(out["iters"][0])
    tot = p[:12].sum() + p[13:16].sum()
    print(f"B={B} instance 0: status {out['status'][0]} iters {it}; total {tot} cycles = {tot / 1.965e6:.3f} ms")
    for k in range(16):
        if p[k]:
            print(f"  {NAMES[k]:42s} {p[k]:10d} cyc {100 * p[k] / tot:5.1f}%  per iteration {p[k] / max(it, 1):9.0f}")
    CN = ["diag factor+invert (warp 0)", "wait barrier 1", "phase 2 (panel DMMA + column j+1)", "wait barrier 2", "16x16 inverses", "32x32 inverses"]
    nfac = it + 1   # factorisations of the last solve: starting point + one per iteration (counters are reset per read)
    for k in range(6):
        print(f"    chol::factor {CN[k]:36s} {p[16 + k] / nfac:9.0f} cyc per factorisation")
    KN = ["tables", "dense part (kkt_dense_mma)", "-", "position x position", "force-sample items + barrier", "  dense: H blocks + first chunk staged", "  dense: chunk loop", "-"]
    for k in range(8):
        print(f"    kkt_assemble {KN[k]:36s} {p[24 + k] / nfac:9.0f} cyc per assembly")
