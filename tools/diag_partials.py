"""Developer diagnostic: entries of the exported parameter partials (bgg_param_partials) whose value or non-zero pattern differs from the
oracle's, for every contact time of one mirrored instance."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from common import wl
import test_gpu_parity as tp
cfg_name = os.environ.get("CFG", "a1_configuration")
cfg = wl.CONFIGS[cfg_name]
states, _, ee = wl.batched_trot_inputs(cfg, 2, seed=21)
states[0] = cfg["srb_init"]; ee[0] = wl.EE_NOMINAL
gpu, oracles, out, go = tp._gradient_case(cfg_name, states, ee)
o = oracles[0]
common.mirror_oracle_to_gpu(o, gpu, 0)
ct = go.contact_times(o)
sz = o.sizes()
print({k: sz[k] for k in ("n", "num_eq", "num_ineq", "num_dyn", "num_force_box", "num_cone")})
for foot in range(4):
    for idx in range(len(ct[foot][0])):
        want = go.param_partials(o, foot, idx); got = gpu.ComputeParamPartialsClarabel(0, foot, idx)
        for key in ("dA", "dG"):
            w = want[key].toarray(); g = got[key]
            bad = np.argwhere((g != 0) != (w != 0))
            if len(bad):
                print(foot, idx, key, "pattern differs at", len(bad), [(int(r), int(c), g[r, c], w[r, c]) for r, c in bad[:6]])
