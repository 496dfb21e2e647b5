import sys
sys.path.insert(0, "tests")
import numpy as np, common
from common import wl
for cfg_name in ("a1_configuration", "a1_gait_opt_config"):
    cfg = wl.CONFIGS[cfg_name]
    init = np.asarray(cfg["srb_init"], float)
    o = common.make_oracle(cfg_name)
    gpu = common.make_gpu(cfg_name, 1)
    ee = wl.EE_NOMINAL.copy()
    o.initial_run(init, ee)
    state = init.copy()
    seen = []
    dt = cfg["integrator_dt"]
    for step in range(40):
        t0 = dt * step
        common.mirror_oracle_to_gpu(o, gpu, 0)
        ee_now = np.array([o.ee_at(e, t0) for e in range(4)])
        out = gpu.GetRealTimeUpdate(state[None], np.array([t0]), ee_now[None])
        st = o.solve(state, t0, ee_now, real_time=True)
        sz = gpu.sizes(0)
        rel = np.linalg.norm(gpu.solution(0)["qp_sol"] - o.qp_solution()["x"]) / np.linalg.norm(o.qp_solution()["x"])
        seen.append((sz["nu"], sz["n_samples"], int(out["status"][0]), int(st), float(f"{rel:.1e}")))
        state = o.states()[1].copy()
    print(cfg_name, sorted(set((a, b) for a, b, _, _, _ in seen)))
    print([s for s in seen if s[2] != 0 or s[3] != 0 or s[4] > 1e-4][:10])
